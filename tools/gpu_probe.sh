L=gpurun_out/r1c_linear_B.log
: > $L
echo "--- pair, residual" >> $L
timeout 300 python tools/bench_linear.py --model B >> $L 2>&1
echo "--- pair, no residual" >> $L
timeout 300 python tools/bench_linear.py --model B --no-res --no-check >> $L 2>&1
echo "--- L pair, residual" >> $L
timeout 300 python tools/bench_linear.py --model L >> $L 2>&1
cat $L
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
