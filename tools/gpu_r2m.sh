mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2m_bench_n8.json").read().strip().splitlines()[-1])
    print("n8 ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), d["config"]["exchange"], "extra", d.get("extra_workload") and (round(d["extra_workload"]["value"],1), round(d["extra_workload"]["ms_per_step"],2)), d["clocks"])
except Exception as e: print("ERR", e)
PY
tail -4 gpurun_out/r2m_bench_n8.err
