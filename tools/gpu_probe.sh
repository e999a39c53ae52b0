# Full regression on one B200 (run with: gpurun --timeout 2400 -- 'bash tools/gpu_probe.sh')
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py 2>&1 | tail -1
