# Round-end regression on one B200: smoke, the whole -m gpu suite, the bench line (both arms), the GEMM capture the bench's
# `roofline.traffic` is read from, and the HBM-bound kernel table.   gpurun --timeout 2400 -- 'bash tools/gpu_final.sh'
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/final_smoke.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/final_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|^E  |FAILED" gpurun_out/final_tests.log | head -10
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
if [ -z "$VPT_FINAL_LIGHT" ]; then   # VPT_FINAL_LIGHT=1: skip the reference arm and the GEMM capture (unchanged kernels)
VPT_CPU_BUDGET_S=40 timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?"
fi
timeout 300 python tools/bench_membound.py > gpurun_out/final_membound.txt 2>&1; echo "membound rc=$?"
[ -z "$VPT_FINAL_LIGHT" ] && python tools/profile_step.py > gpurun_out/final_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair -s 20 -c 8 -o gpurun_out/final_gemm python tools/profile_step.py > gpurun_out/final_ncu.log 2>&1; echo "gemm capture rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_step_launches.csv python tools/profile_step.py > gpurun_out/final_launches.log 2>&1; echo "launch list rc=$?"
python tools/summarize_launches.py gpurun_out/final_step_launches.csv > gpurun_out/final_step_launches_summary.txt 2>&1; head -12 gpurun_out/final_step_launches_summary.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/final_bench.json"))
print("ms/step", round(d["ms_per_step"],3), "img/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3), round(d["roofline"]["gemm_only"]["frac"],3), "launches", d["gpu_launches_per_step"], d["clocks"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "| ref gpu", d["reference_gpu"].get("value"), d["vs_reference_gpu"], "| JiT-L", d["extra_workload"]["value"], d["extra_workload"]["ms_per_step"])
import os
if not os.environ.get("VPT_FINAL_LIGHT"):
    r=json.load(open("gpurun_out/final_bench_reference.json"))
    print("reference arm", r["value"], r["steps"], r["cpu_baseline"]["kind"], r["wall_s"])
PY
