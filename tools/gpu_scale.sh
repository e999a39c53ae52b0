# Scaling check on one 8-GPU box (charged 8x):  gpurun --gpus 8 --timeout 700 -- 'bash tools/gpu_scale.sh'
mkdir -p gpurun_out
for n in 8 4; do
  timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "bench n=$n rc=$?"
done
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f"gpurun_out/scale_n{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"], 3), "img/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "JiT-L", round(d["extra_workload"]["value"], 1), round(d["extra_workload"]["ms_per_step"], 2), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "ERR", e)
PY
