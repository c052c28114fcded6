set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | wc -l
timeout 200 $TR --nproc-per-node 8 --master-port 29511 scripts/check_slab.py --shape 64 64 128 --steps 7 --transport fused 2>&1 | grep -E "SLAB|Error|error|Traceback" -A3 | head -20
timeout 240 $TR --nproc-per-node 8 --master-port 29513 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport fused --time-steps 300 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --nproc-per-node 8 --master-port 29514 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport nccl --time-steps 300 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --nproc-per-node 4 --master-port 29515 scripts/check_slab.py --shape 512 512 512 --steps 5 --transport fused --time-steps 300 2>&1 | grep -E "SLAB|Error|error" | head
timeout 240 $TR --nproc-per-node 8 --master-port 29516 scripts/check_slab.py --shape 128 128 128 --steps 5 --transport fused --time-steps 500 2>&1 | grep -E "SLAB|Error|error" | head
for n in 8 4; do
timeout 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_r01_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ['value','n_gpus','ms_per_step','gpu_launches']}, d['roofline']['frac'], d['e2e']['value'], d['halo']['transport'], d['clocks'])"
done
