timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== fused monomial sums"; timeout 200 python scripts/perf_bwd.py
echo "== separate monomial kernel"; PERCNN_BWD_SPLIT_MONO=1 timeout 200 python scripts/perf_bwd.py 2>&1 | grep 512
