set -x
timeout 180 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 || { echo SMOKE_FAILED; exit 1; }
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 500 python scripts/perf_cfgs.py
