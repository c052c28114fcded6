echo "== PERCNN_BWD_MW=1 tests"; PERCNN_BWD_MW=1 timeout 200 python -m pytest tests -m gpu -q -x -k "gs3d or tma or data_loss or lattice" 2>&1 | tail -4
echo "== MW perf"; PERCNN_BWD_MW=1 timeout 100 python scripts/perf_bwd.py 2>&1 | tail -6
echo "== MW no-mono"; PERCNN_BWD_MW=1 PERCNN_KERNEL_DEBUG=8 timeout 100 python scripts/perf_bwd.py 2>&1 | tail -2
