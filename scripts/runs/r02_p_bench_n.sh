mkdir -p gpurun_out
N=${N:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; tail -2 gpurun_out/r02_bench_n$N.err | cut -c1-300
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n$N.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
for k in ('halo_check', 'train_gs3d_512', 'cfg4_gs3d_128'):
    v = d.get(k)
    if isinstance(v, dict): v = {a: b for a, b in v.items() if a not in ('note', 'includes', 'halo', 'against')}
    print(k, v)
PY
