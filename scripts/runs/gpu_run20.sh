NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for pdl in 0 1; do
PERCNN_NO_PDL=$pdl timeout 200 $TR --master-port 2955$pdl scripts/check_slab.py --shape 128 512 512 --steps 25 --repeat 6 --transport fused --time-steps 400 2>&1 | grep -E "^SLAB|Error|error" | head -4
PERCNN_NO_PDL=$pdl timeout 200 $TR --master-port 2956$pdl scripts/check_slab.py --shape 48 64 128 --steps 40 --repeat 10 --transport fused 2>&1 | grep -E "^SLAB|Error|error" | head -4
done
timeout 200 $TR --master-port 29571 scripts/check_slab_bwd.py --shape 64 48 128 --steps 6 2>&1 | grep -E "^SLAB|Error|error" | head -4
