set -x
NG=${NG:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29511 scripts/check_slab.py --shape 64 64 128 --steps 7 --transport nccl 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --master-port 29512 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport nccl --time-steps 200 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --master-port 29513 scripts/check_slab.py --shape 64 64 128 --steps 7 --transport symm 2>&1 | grep -E "SLAB|Error|error|Traceback" -A3 | head -20
timeout 240 $TR --master-port 29514 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport symm --time-steps 200 2>&1 | grep -E "SLAB|Error|error" | head
timeout 400 $TR --master-port 29515 bench.py --gpus $NG --steps 3 --warmup 3 2>&1 | tail -2
nvidia-smi topo -m | head -12
