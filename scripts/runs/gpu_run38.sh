mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r01i_pytest.log; tail -3 gpurun_out/r01i_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01i_bench_ref.json 2>gpurun_out/r01i_bench_ref.err; cut -c1-200 gpurun_out/r01i_bench_ref.json
timeout 600 python bench.py > gpurun_out/r01i_bench_n1.json 2> gpurun_out/r01i_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01i_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d.get('train_gs3d_512',{}).get('ms_per_timestep_fwd_plus_adjoint'), d['cpu_baseline']['value'])"; tail -3 gpurun_out/r01i_bench_n1.err
timeout 400 python scripts/perf_cfgs.py > gpurun_out/r01i_perf_cfgs.txt 2>&1; tail -20 gpurun_out/r01i_perf_cfgs.txt
timeout 60 python scripts/perf_fwd.py 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01i_launches_512.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01i_ncu_bench.log 2>&1; tail -c 300 gpurun_out/r01i_ncu_bench.log; wc -l gpurun_out/r01i_launches_512.csv
