mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r01f_pytest.log; tail -8 gpurun_out/r01f_pytest.log
timeout 300 python scripts/perf_cfg1_breakdown.py 2>&1 | tail -9
timeout 400 python scripts/perf_cfgs.py > gpurun_out/r01f_perf_cfgs.txt 2>&1; cat gpurun_out/r01f_perf_cfgs.txt
