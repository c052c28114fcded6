set -x
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 200 $TR --master-port 29511 scripts/check_slab.py --shape 64 64 128 --steps 7 --transport fused 2>&1 | grep -E "SLAB|Error|error|Traceback" -A3 | head -20
timeout 200 $TR --master-port 29512 scripts/check_slab.py --shape 40 48 256 --steps 9 --transport fused 2>&1 | grep -E "SLAB|Error|error|Traceback" -A3 | head -20
timeout 240 $TR --master-port 29513 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport fused --time-steps 200 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --master-port 29514 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport symm --time-steps 200 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --master-port 29516 scripts/check_slab.py --shape 128 128 128 --steps 5 --transport fused --time-steps 500 2>&1 | grep -E "SLAB|Error|error" | head
timeout 400 $TR --master-port 29515 bench.py --gpus $NG --steps 3 --warmup 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches']}, d['roofline']['frac'], d['e2e']['value'], d['halo'])"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches']}, d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
