echo "== SDXL D=640 (L=4096, B=4): linears"; timeout 300 python tools/bench_linear.py --model sdxl640 --M 16384 --no-res 2>&1 | tail -9 | cut -c1-110
echo "== SDXL D=1280 (L=1024, B=4): linears"; timeout 300 python tools/bench_linear.py --model sdxl1280 --M 4096 --no-res 2>&1 | tail -9 | cut -c1-110
echo "== attention"; timeout 200 python tools/bench_attn.py --B 4 --H 10 --L 4096 2>&1 | tail -2
timeout 200 python tools/bench_attn.py --B 4 --H 10 --L 4096 --Lk 231 2>&1 | tail -2
timeout 200 python tools/bench_attn.py --B 4 --H 20 --L 1024 2>&1 | tail -2
timeout 200 python tools/bench_attn.py --B 4 --H 20 --L 1024 --Lk 231 2>&1 | tail -2
