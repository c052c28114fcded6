timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
exec(open("scripts/sweep_tma.py").read().split("configs = [")[0])
for name, env in [("default (PDL)", {}), ("no PDL", {"PERCNN_NO_PDL": "1"}), ("N=256 PDL", {"N": "256", "STEPS": "200"}), ("N=256 no PDL", {"N": "256", "STEPS": "200", "PERCNN_NO_PDL": "1"}),
                  ("N=128 PDL", {"N": "128", "STEPS": "500"}), ("N=128 no PDL", {"N": "128", "STEPS": "500", "PERCNN_NO_PDL": "1"})]:
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "-c", CHILD], env=e, capture_output=True, text=True, timeout=120)
    print(f"{name:28s} {out.stdout.strip() or out.stderr.strip()[-300:]}", flush=True)
PY
timeout 200 python scripts/perf_bwd.py 2>&1 | grep "g_add=True"
