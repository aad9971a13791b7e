# development aid: host-call (e2e) leg at N ranks for several shard plans (lanes first_mb growth)
N=${1:-8}
for cfg in "1 10 3" "1 15 2.34" "1 6 4"; do set -- $cfg; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --e2e-only --lanes $1 --plan-first-mb $2 --plan-growth $3 2>/dev/null | grep e2e_only; done
