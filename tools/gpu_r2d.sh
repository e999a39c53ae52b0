mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_block.py -q -x > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2d_tests.log
VPT_FUSE_SWIGLU=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2d_bench_unfused.json 2> gpurun_out/r2d_bench_unfused.err; echo "bench0 rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2d_bench_fused.json 2> gpurun_out/r2d_bench_fused.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in ("unfused","fused"):
    try:
        d=json.load(open(f"gpurun_out/r2d_bench_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), round(d["roofline"]["gemm_only"]["frac"],3), "launches", d["gpu_launches_per_step"], d["clocks"])
    except Exception as e: print(n, "ERR", e)
PY
tail -3 gpurun_out/r2d_bench_fused.err
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_train.py -q -s 2>&1 | grep -E "^\[|passed|failed|checkpointing" 
