set -x
timeout 180 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 || { echo SMOKE_FAILED; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python scripts/sweep_tma.py
