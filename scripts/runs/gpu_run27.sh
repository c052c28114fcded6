timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/perf_cfgs.py 2>&1 | grep -E "cfg1|cfg2|cfg3ii|ref-size|GS2D"
