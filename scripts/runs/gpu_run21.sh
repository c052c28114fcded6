NG=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
PERCNN_NO_PDL=1 timeout 120 python scripts/check_determinism.py
PERCNN_NO_PDL=0 timeout 120 python scripts/check_determinism.py
PERCNN_NO_PDL=1 timeout 200 $TR --master-port 29551 scripts/check_slab.py --shape 128 512 512 --steps 25 --repeat 6 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | head -12
PERCNN_NO_PDL=1 timeout 200 $TR --master-port 29552 scripts/check_slab.py --shape 128 512 512 --steps 25 --repeat 6 --transport symm 2>&1 | grep -E "^SLAB|MISMATCH" | head -12
PERCNN_NO_PDL=0 timeout 200 $TR --master-port 29553 scripts/check_slab.py --shape 128 512 512 --steps 25 --repeat 12 --transport fused 2>&1 | grep -E "^SLAB|MISMATCH" | head -12
