mkdir -p gpurun_out
timeout 600 python tools/probes/determinism.py > gpurun_out/r2c_determinism.log 2>&1; echo "det rc=$?"
timeout 600 python tools/probes/e2e_variants.py > gpurun_out/r2c_e2e.log 2>&1; echo "e2e rc=$?"
tail -30 gpurun_out/r2c_determinism.log; cat gpurun_out/r2c_e2e.log | tail -12
