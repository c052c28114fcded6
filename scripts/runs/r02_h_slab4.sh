mkdir -p gpurun_out
N=${N:-4}
for shape in "512 512 512" "128 128 128"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/check_slab.py --shape $shape --steps 9 --transport fused --time-steps 500 2>&1 | grep -E "SLAB_|MISMATCH|Error|error" | head -8
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29703 scripts/check_slab_bwd.py --shape 64 48 128 --steps 6 2>&1 | grep -E "SLAB_|Error|error" | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29704 scripts/check_slab.py --shape 128 512 512 --steps 5 --repeat 20 --transport fused 2>&1 | grep -E "SLAB_|MISMATCH|Error|error" | head -8
