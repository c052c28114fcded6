mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r01d_pytest.log; tail -8 gpurun_out/r01d_pytest.log
timeout 600 python bench.py > gpurun_out/r01d_bench_n1.json 2> gpurun_out/r01d_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r01d_bench_n1.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d.get('train_gs3d_512'))"; tail -3 gpurun_out/r01d_bench_n1.err
timeout 400 python scripts/perf_cfgs.py > gpurun_out/r01d_perf_cfgs.txt 2>&1; cat gpurun_out/r01d_perf_cfgs.txt
