timeout 300 python tools/profile_step.py > gpurun_out/r1n_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1n_step_launches.csv python tools/profile_step.py > gpurun_out/r1n_ncu.log 2>&1
tail -2 gpurun_out/r1n_ncu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['clocks'])"
