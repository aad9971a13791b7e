# development aid: host-call (e2e) leg at N ranks for several shard plans
N=${1:-8}
for cfg in "2 50 1" "2 0 0" "2 12 2.5" "2 25 2" "3 12 2.5" "1 25 3"; do set -- $cfg; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --e2e-only --lanes $1 --plan-first-mb $2 --plan-growth $3 2>/dev/null | grep e2e_only; done
