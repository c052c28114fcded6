echo "== fwd"; timeout 60 python scripts/perf_fwd.py 2>&1 | tail -2
echo "== bwd"; timeout 60 python scripts/perf_bwd.py 2>&1 | tail -2
timeout 300 python -m pytest tests -m gpu -q -x -k "gs3d or tma or data_loss or lattice" 2>&1 | tail -2
