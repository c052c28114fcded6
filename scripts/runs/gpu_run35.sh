mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== with PERCNN_BWD_MW=1 (experimental kernel)"; PERCNN_BWD_MW=1 timeout 300 python -m pytest tests -m gpu -q -k "gs3d or tma or data_loss or lattice" 2>&1 | tail -3
timeout 100 python scripts/perf_bwd.py 2>&1 | tail -2
