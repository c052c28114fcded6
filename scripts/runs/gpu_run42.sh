mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/r01j_bench_n1.json 2> gpurun_out/r01j_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01j_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d.get('train_gs3d_512',{}).get('ms_per_timestep_fwd_plus_adjoint'), d['cpu_baseline']['value'])"; tail -3 gpurun_out/r01j_bench_n1.err
