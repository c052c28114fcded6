mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r01c_pytest.log; tail -6 gpurun_out/r01c_pytest.log
timeout 200 python scripts/perf_bwd.py > gpurun_out/r01c_perf_bwd.txt 2>&1; cat gpurun_out/r01c_perf_bwd.txt
