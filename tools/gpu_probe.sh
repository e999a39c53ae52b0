timeout 300 python tools/profile_step.py --model JiT-H/16 --res 512 --batch 16 > gpurun_out/r1s_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_kernel|attn_bwd2_kernel" -s 4 -c 1 -o gpurun_out/r1s_attn80_fwd python tools/profile_step.py --model JiT-H/16 --res 512 --batch 16 > gpurun_out/r1s_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_bwd2_kernel" -s 2 -c 1 -o gpurun_out/r1s_attn80_bwd python tools/profile_step.py --model JiT-H/16 --res 512 --batch 16 > gpurun_out/r1s_ncu2.log 2>&1
ls -la gpurun_out/r1s*
