mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dp.py -q -s > gpurun_out/r2f_dp_tests.log 2>&1; echo "dp tests rc=$?"
grep -E "DP world|passed|failed|Error" gpurun_out/r2f_dp_tests.log | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err; echo "bench n2 rc=$?"
VPT_NCCL_IN_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r2f_bench_n2_nograph.json 2> gpurun_out/r2f_bench_n2_nograph.err; echo "bench n2 (exchange outside graph) rc=$?"
VPT_DP_CHUNKS=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r2f_bench_n2_1chunk.json 2> gpurun_out/r2f_bench_n2_1chunk.err; echo "bench n2 (1 chunk) rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for n in ("n1","n2","n2_nograph","n2_1chunk"):
    try:
        d=json.loads(open(f"gpurun_out/r2f_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["exchange"], "extra", d.get("extra_workload") and (round(d["extra_workload"]["value"],1), round(d["extra_workload"]["ms_per_step"],2)))
    except Exception as e: print(n, "ERR", e)
PY
tail -5 gpurun_out/r2f_bench_n2.err
