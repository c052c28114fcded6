mkdir -p gpurun_out
# full GPU test suite
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_final_pytest_gpu.log
# every launch of a short default bench run with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench_n1.log 2>&1
tail -2 gpurun_out/r02_launches_bench_n1.log | cut -c1-200
# full captures of the flagship kernels
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_gs3d|k_tile2d|k_pi_k5|k_multi_step_slab_tb" -o gpurun_out/r02_ncu_kernels -f python scripts/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1
tail -2 gpurun_out/r02_ncu_kernels.log
ls -la gpurun_out/r02_ncu_kernels.ncu-rep
# the real bench line (not under a profiler)
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python scripts/perf_cfgs.py > gpurun_out/r02_perf_all_configs.txt 2>&1; cat gpurun_out/r02_perf_all_configs.txt
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_bench_n1.json') if l.startswith('{')][-1])
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
print('train', {k: v for k, v in d['train_gs3d_512'].items() if k != 'note'})
for k, v in d['configs'].items():
    print(k, {a: b for a, b in v.items() if a not in ('workload',)})
PY
