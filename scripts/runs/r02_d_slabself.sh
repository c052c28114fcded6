mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_self_gpu.py -q -x 2>&1 | tail -15
timeout 600 python scripts/perf_slab_self.py 2>&1 | tail -12 | tee gpurun_out/r02d_perf_slab_self.txt
