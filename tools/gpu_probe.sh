timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -1 > gpurun_out/r1o_bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/r1o_bench_n2.json')); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d.get('extra_workload'))"
