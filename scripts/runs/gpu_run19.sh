NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29541 scripts/check_slab.py --shape 128 512 512 --steps 5 --transport fused --time-steps 400 2>&1 | grep -E "^SLAB|Error|error" | head
PERCNN_NO_PDL=1 timeout 200 $TR --master-port 29542 scripts/check_slab.py --shape 128 512 512 --steps 5 --transport fused --time-steps 400 2>&1 | grep -E "^SLAB|Error|error" | head
timeout 200 $TR --master-port 29543 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport fused --time-steps 200 2>&1 | grep -E "^SLAB|Error|error" | head
timeout 100 python - <<'PY'
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
exec(open("scripts/sweep_tma.py").read().split("configs = [")[0])
CH = CHILD.replace('shape = (n, n, n)', 'shape = (64, 512, 512)').replace("n**3*16", "64*512*512*16").replace("synthetic_state(shape, 0, n,", "synthetic_state(shape, 0, 64,")
out = subprocess.run([sys.executable, "-c", CH], env=dict(os.environ, N="512", STEPS="400"), capture_output=True, text=True, timeout=90)
print("single GPU 64x512x512 periodic:", out.stdout.strip() or out.stderr.strip()[-300:], flush=True)
PY
