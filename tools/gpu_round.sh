set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1b_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1b_gputests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; echo "bench rc=$?"
python tools/profile_step.py > gpurun_out/r1b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_nf4lora -s 20 -c 4 -o gpurun_out/r1b_gemm python tools/profile_step.py > gpurun_out/r1b_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"attn_fwd_kernel|attn_bwd_kernel|lora_grad_kernel|rmsnorm_bwd" -s 8 -c 6 -o gpurun_out/r1b_attn python tools/profile_step.py > gpurun_out/r1b_ncu_attn.log 2>&1
tail -3 gpurun_out/r1b_gputests.log; cat gpurun_out/r1b_bench.json
