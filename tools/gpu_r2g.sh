mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -q -s > gpurun_out/r2g_dp_tests.log 2>&1; echo "dp tests rc=$?"
grep -E "DP world|passed|failed|Error|assert" gpurun_out/r2g_dp_tests.log | head
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for n in ("n2",):
    try:
        d=json.loads(open(f"gpurun_out/r2g_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["exchange"])
    except Exception as e: print(n, "ERR", e)
PY
tail -3 gpurun_out/r2g_bench_n2.err
