python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r1f_launches.csv python tools/profile_step.py > gpurun_out/ncu_f.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
