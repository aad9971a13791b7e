# development aid: round-0 group size sweep (device-resident compress, 100 MB level 9)
for t in 100000 3552 1776 888 444 220; do echo "R0_TILES=$t"; BZ2B200_R0_TILES=$t BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 6 --warmup 3; done
BZ2B200_R0_TILES=888 BZ2B200_TRACE=1 BENCH_NO_SAMPLER=1 python bench.py --profile-only --steps 1 --warmup 3 2>&1 | grep "trace" | tail -45
