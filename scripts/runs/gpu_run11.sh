set -x
timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python scripts/perf_cfgs.py 2>&1 | grep -E "GS3D|cfg4"
