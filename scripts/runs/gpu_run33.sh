mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q -x -k "gs3d or tma or data_loss or lattice" 2>&1 | tail -15 > gpurun_out/r01g_pytest_tma.log; tail -6 gpurun_out/r01g_pytest_tma.log
echo "== MW kernel"; timeout 100 python scripts/perf_bwd.py 2>&1 | tail -6
echo "== old kernel"; PERCNN_BWD_MW=0 timeout 100 python scripts/perf_bwd.py 2>&1 | tail -6
