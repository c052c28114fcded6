timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/perf_bwd.py
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches']}, d['roofline']['frac'], d['e2e'], d['clocks'])"
