mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r02j_bench_n8.json 2> gpurun_out/r02j_bench_n8.err; tail -3 gpurun_out/r02j_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02j_bench_n8.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'], d['step_ms'])
for k in ('halo_check', 'train_gs3d_512', 'cfg4_gs3d_128'):
    v = d.get(k); 
    if isinstance(v, dict): v = {a: b for a, b in v.items() if a not in ('note', 'includes', 'halo', 'against')}
    print(k, v)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/check_slab.py --shape 128 128 128 --steps 9 --transport fused --time-steps 500 2>&1 | grep -E "SLAB_|MISMATCH|Error|error" | head -4
PERCNN_SLAB_NO_PERSISTENT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29705 scripts/check_slab.py --shape 128 128 128 --steps 9 --transport fused --time-steps 500 2>&1 | grep -E "SLAB_|MISMATCH|Error|error" | head -4
