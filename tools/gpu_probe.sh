python tools/profile_step.py > gpurun_out/plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --profile-from-start off -k regex:gemm_pair --csv --log-file gpurun_out/r1g_gemm_pair_metrics.csv python tools/profile_step.py > gpurun_out/ncu_g.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_pair -s 30 -c 3 -o gpurun_out/r1g_gemm_pair python tools/profile_step.py > gpurun_out/ncu_g2.log 2>&1
tail -2 gpurun_out/ncu_g.log gpurun_out/ncu_g2.log
