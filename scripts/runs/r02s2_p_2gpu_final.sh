mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "training or forward_is_bitwise" 2>&1 | tail -2
N=2 bash scripts/runs/r02_p_bench_n.sh
