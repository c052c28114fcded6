mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_slab_self_gpu.py -x -q -m gpu 2>&1 | tail -2
N=2 bash scripts/runs/r02_p_bench_n.sh
