# development aid: time every build in gpurun_variants/ (device-resident compress; the scatter passes from the trace)
echo "== default build, BZ2B200_RS2=0 (first-generation passes)"
BZ2B200_RS2=0 BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 12 --warmup 3
for f in gpurun_variants/lib_*.so; do
  echo "== $f"
  BZ2B200_LIB=$PWD/$f BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 12 --warmup 3
  BZ2B200_LIB=$PWD/$f BZ2B200_TRACE=1 BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 1 --warmup 3 2>&1 | grep "trace" | tail -44 | grep -E "launches|k_rs_scatter2|k_rs_bases|k_rs_scan<9>"
done
