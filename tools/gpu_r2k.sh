mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/r2k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2k_step_launches.csv python tools/profile_step.py > gpurun_out/r2k_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair -s 20 -c 8 -o gpurun_out/r2k_gemm python tools/profile_step.py > gpurun_out/r2k_ncu2.log 2>&1; echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_kernel|attn_bwd2_kernel|lora_grad_kernel" -s 6 -c 3 -o gpurun_out/r2k_attn python tools/profile_step.py > gpurun_out/r2k_ncu3.log 2>&1; echo "attn full rc=$?"
python tools/bench_membound.py --once > gpurun_out/r2k_mb_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2k_membound.csv python tools/bench_membound.py --once > gpurun_out/r2k_ncu4.log 2>&1; echo "membound counters rc=$?"
python bench.py --steps 2 --warmup 3 --settle-s 0 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2k_bench_plain.json 2> gpurun_out/r2k_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2k_bench_launches.csv python bench.py --steps 2 --warmup 3 --settle-s 0 --no-cpu-baseline --no-reference-gpu --no-extra > gpurun_out/r2k_ncu5.log 2>&1; echo "bench launch list rc=$?"
ls -la gpurun_out/r2k_*
python tools/summarize_launches.py gpurun_out/r2k_step_launches.csv | head -24
