mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.log; tail -8 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02a_bench_n1.json 2> gpurun_out/r02a_bench_n1.err; tail -3 gpurun_out/r02a_bench_n1.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02a_bench_n1.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
print('train', d.get('train_gs3d_512'))
for k, v in d.get('configs', {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != 'workload'})
print('cpu', d.get('cpu_baseline'))
PY
