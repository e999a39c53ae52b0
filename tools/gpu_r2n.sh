mkdir -p gpurun_out
timeout 500 python bench.py --model JiT-H/16 --res 512 --batch 16 --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2n_jith_bench.json 2> gpurun_out/r2n_jith_bench.err; echo "jit-h bench rc=$?"
timeout 500 python tools/bench_arb.py --steps 27 > gpurun_out/r2n_jith_arb.json 2> gpurun_out/r2n_jith_arb.err; echo "jit-h arb rc=$?"
timeout 400 python bench.py --checkpointing --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2n_jitb_ckpt.json 2> gpurun_out/r2n_jitb_ckpt.err; echo "ckpt bench rc=$?"
timeout 400 python bench.py --optimizer radam_schedulefree --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2n_jitb_sf.json 2> gpurun_out/r2n_jitb_sf.err; echo "schedulefree bench rc=$?"
python - <<'PY'
import json
for n in ("jith_bench","jith_arb","jitb_ckpt","jitb_sf"):
    try:
        d=json.loads(open(f"gpurun_out/r2n_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", d.get("e2e",{}).get("value"), "roof", d.get("roofline") and round(d["roofline"]["frac"],3), "tflops", d.get("achieved_tflops_step"), d.get("ms_per_step_by_bucket"))
    except Exception as e: print(n, "ERR", e)
PY
