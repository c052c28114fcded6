timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/perf_cfgs.py
echo "== without the persistent kernel"; PERCNN_NO_MULTISTEP=1 timeout 300 python scripts/perf_cfgs.py 2>&1 | grep -E "cfg1|cfg2|cfg3ii|ref-size|GS2D"
