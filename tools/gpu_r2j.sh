mkdir -p gpurun_out
timeout 500 python tools/ab_step.py > gpurun_out/r2j_ab_jitb.txt 2>&1; echo "ab rc=$?"; cat gpurun_out/r2j_ab_jitb.txt | tail -5
timeout 500 python tools/ab_step.py --model JiT-L/16 --rounds 3 --steps 10 > gpurun_out/r2j_ab_jitl.txt 2>&1; echo "ab-L rc=$?"; cat gpurun_out/r2j_ab_jitl.txt | tail -5
