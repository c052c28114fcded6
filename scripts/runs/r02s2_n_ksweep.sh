mkdir -p gpurun_out
N=${N:-8}
O=gpurun_out/r02s2n_cfg4_n${N}.txt
: > $O
for k in $KS; do
PERCNN_SLAB_TB_K=$k timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29620 scripts/check_slab.py --shape 128 128 128 --steps 21 --transport fused --time-steps 500 2>&1 | grep -E "SLAB_" | sed "s/^/K=$k /" >> $O
done
cat $O
