# round-2 first probe: new tests, bench, ATen profile
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dp.py tests/test_gpu_train.py -q -s > gpurun_out/r2a_newtests.log 2>&1; echo "newtests rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_dp.py --deselect tests/test_gpu_train.py > gpurun_out/r2a_oldtests.log 2>&1; echo "oldtests rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 300 python tools/profile_aten.py > gpurun_out/r2a_aten.log 2>&1; echo "aten rc=$?"
tail -5 gpurun_out/r2a_newtests.log; tail -3 gpurun_out/r2a_oldtests.log; cat gpurun_out/r2a_bench.json | head -c 3000
