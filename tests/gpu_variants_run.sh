# development aid: time every build in gpurun_variants/ (device-resident compress; the scatter passes from the trace)
for f in gpurun_variants/lib_*.so; do
  echo "== $f"
  BZ2B200_LIB=$PWD/$f BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 6 --warmup 3
  BZ2B200_LIB=$PWD/$f BZ2B200_TRACE=1 BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 1 --warmup 3 2>&1 | grep "trace" | tail -44 | grep -E "launches|k_rs_scatter2|k_rs_bases|k_sort_groups|k_rank0|k_mtf_ranks"
done
