set -x
timeout 180 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -3 || { echo SMOKE_FAILED; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ['value','ms_per_step','gpu_launches','clocks']}, d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value'], d.get('cfg4_gs3d_128'))"
