mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py -q -x -k "swiglu_epilogues" > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2e_tests.log
VPT_FUSE_SWIGLU=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2e_bench_unfused.json 2> gpurun_out/r2e_bench_unfused.err; echo "bench0 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2e_bench_fused.json 2> gpurun_out/r2e_bench_fused.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in ("unfused","fused"):
    try:
        d=json.load(open(f"gpurun_out/r2e_bench_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), round(d["roofline"]["gemm_only"]["frac"],3), "launches", d["gpu_launches_per_step"], d["clocks"])
    except Exception as e: print(n, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_step_launches.csv python tools/profile_step.py > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/r2e_step_launches.csv > gpurun_out/r2e_step_launches_summary.txt 2>&1; head -30 gpurun_out/r2e_step_launches_summary.txt
