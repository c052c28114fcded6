mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_self_gpu.py -q -x 2>&1 | tail -5
STEPS=300 timeout 300 python scripts/perf_slab_self.py 2>&1 | grep -E "slab-self" | tee gpurun_out/r02i_small.txt
