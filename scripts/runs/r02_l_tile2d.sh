mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tile2d_gpu.py -q -x 2>&1 | tail -8
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 300 python scripts/perf_cfgs.py 2>&1 | tee gpurun_out/r02l_perf_cfgs.txt | grep -E "cfg|ref-size"
