timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_arb.py --steps 18 2>&1 | tail -1 | cut -c1-700
