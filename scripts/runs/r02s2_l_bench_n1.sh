mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
print('train', {k: v for k, v in d['train_gs3d_512'].items() if k != 'note'})
for k, v in d['configs'].items():
    print(k, {a: b for a, b in v.items() if a not in ('workload', 'cpu_baseline')})
PY
PERF_ONLY=cfg3ii python scripts/perf_cfgs.py
