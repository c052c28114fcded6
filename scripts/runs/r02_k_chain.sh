mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_slab_self_gpu.py -q -x 2>&1 | tail -3
for v in "" "PERCNN_FUSED_DEBUG=18" "PERCNN_FUSED_DEBUG=3"; do
echo "== $v"
env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/check_slab.py --shape 128 512 512 --steps 3 --transport fused --time-steps 600 2>&1 | grep -E "SLAB_TIME|SLAB_CHECK" | head -4
done 2>&1 | tee gpurun_out/r02k_chain.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29704 scripts/check_slab.py --shape 128 512 512 --steps 5 --repeat 30 --transport fused 2>&1 | grep -E "SLAB_|MISMATCH" | head -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29703 scripts/check_slab_bwd.py --shape 64 48 128 --steps 6 2>&1 | grep -E "SLAB_" | head
