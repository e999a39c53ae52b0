timeout 300 python tools/profile_step.py > gpurun_out/r1p_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_bwd2_kernel" -s 1 -c 1 -o gpurun_out/r1p_attn_bwd python tools/profile_step.py > gpurun_out/r1p_ncu_bwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_pair_kernel" -s 30 -c 3 -o gpurun_out/r1p_gemm python tools/profile_step.py > gpurun_out/r1p_ncu_gemm.log 2>&1
ls -la gpurun_out/r1p*
