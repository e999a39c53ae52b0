mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2i_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|^E  |Error|FAILED" gpurun_out/r2i_tests.log | head -20
VPT_PREFETCH_DEQUANT=0 timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2i_bench_nopf.json 2> gpurun_out/r2i_bench_nopf.err; echo "bench0 rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2i_bench_pf.json 2> gpurun_out/r2i_bench_pf.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in ("nopf","pf"):
    try:
        d=json.load(open(f"gpurun_out/r2i_bench_{n}.json"))
        print(n, "ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), round(d["roofline"]["gemm_only"]["frac"],3), "launches", d["gpu_launches_per_step"], d["clocks"]["sm_mhz"])
    except Exception as e: print(n, "ERR", e)
PY
tail -3 gpurun_out/r2i_bench_pf.err
timeout 300 python tools/bench_membound.py 2>&1 | grep -E "ln_modulate|layernorm|gated_act" 
