# development aid: first- vs second-generation scatter passes (device-resident compress, 100 MB level 9)
python -m pytest tests/test_gpu_parity.py -x -q -k "edge or samples or adversarial or stage or full_size" 2>&1 | tail -3
for v in 0 1; do echo "RS2=$v"; BZ2B200_RS2=$v BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 6 --warmup 3; done
BZ2B200_TRACE=1 BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 1 --warmup 3 2>&1 | grep "trace" | tail -44 | head -16
