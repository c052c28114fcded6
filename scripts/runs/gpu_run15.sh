NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29531 scripts/check_slab_bwd.py --shape 64 48 128 --steps 6 2>&1 | grep -E "SLAB|Error|error|Traceback" -A4 | head -30
timeout 200 $TR --master-port 29532 scripts/check_slab_bwd.py --shape 128 128 128 --steps 10 2>&1 | grep -E "SLAB|Error|error|Traceback" -A4 | head -30
