NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
for st in 1 2 3; do
timeout 250 $TR --master-port 2956$st scripts/check_slab.py --shape 128 512 512 --steps $st --repeat 150 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | head -8
done
