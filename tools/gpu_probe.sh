set -x
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1o_plain.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r1o_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1o_ncu_list.log 2>&1
tail -2 gpurun_out/r1o_ncu_list.log
timeout 300 python tools/profile_step.py > gpurun_out/r1o_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_kernel|attn_bwd2_kernel" -s 2 -c 2 -o gpurun_out/r1o_attn python tools/profile_step.py > gpurun_out/r1o_ncu_attn.log 2>&1
tail -2 gpurun_out/r1o_ncu_attn.log
ls -la gpurun_out/r1o*
