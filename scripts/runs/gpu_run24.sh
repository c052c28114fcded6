NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 250 $TR --master-port 29581 scripts/check_slab.py --shape 128 512 512 --steps 1 --repeat 300 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | cut -c1-160 | tail -4
timeout 250 $TR --master-port 29582 scripts/check_slab.py --shape 128 512 512 --steps 3 --repeat 200 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | cut -c1-160 | tail -4
timeout 250 $TR --master-port 29583 scripts/check_slab.py --shape 128 512 512 --steps 40 --repeat 10 --transport fused --time-steps 400 2>&1 | grep -E "^SLAB|MISMATCH" | cut -c1-160 | tail -4
PERCNN_NO_PDL=1 timeout 250 $TR --master-port 29584 scripts/check_slab.py --shape 128 512 512 --steps 25 --repeat 12 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | cut -c1-160 | tail -4
timeout 250 $TR --master-port 29585 scripts/check_slab.py --shape 48 64 128 --steps 40 --repeat 20 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | cut -c1-160 | tail -4
timeout 200 $TR --master-port 29586 scripts/check_slab_bwd.py --shape 128 128 128 --steps 10 2>&1 | grep -E "^SLAB|Error|error" | head -4
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
