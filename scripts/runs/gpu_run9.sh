set -x
timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
exec(open("scripts/sweep_tma.py").read().split("configs = [")[0])
import subprocess
for name, env in [("default", {}), ("N=256", {"N": "256", "STEPS": "200"}), ("N=128", {"N": "128", "STEPS": "500"})]:
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=120)
    print(f"{name:28s} {out.stdout.strip() or out.stderr.strip()[-300:]}", flush=True)
PY
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches']}, d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
