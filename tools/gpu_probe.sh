# Full regression on one B200:  gpurun --timeout 2600 -- 'bash tools/gpu_final.sh'
bash tools/gpu_final.sh
