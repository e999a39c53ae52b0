timeout 900 python -m pytest tests/test_gpu_attention.py tests/test_gpu_block.py -x -q 2>&1 | tail -2
timeout 120 python tools/probes/small_kernels.py && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"delta" --csv python tools/probes/small_kernels.py 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
for r in rows[1:]:
    if len(r)>14: print(r[4][:40], r[12], r[14])"
timeout 200 python tools/bench_attn.py 2>&1 | tail -1
