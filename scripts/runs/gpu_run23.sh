NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for dbg in 0 1 2 4 7; do
echo "== PERCNN_FUSED_DEBUG=$dbg"
PERCNN_FUSED_DEBUG=$dbg timeout 250 $TR --master-port 2957$dbg scripts/check_slab.py --shape 128 512 512 --steps 1 --repeat 200 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | cut -c1-150 | tail -4
done
