mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02n_pytest.log; tail -40 gpurun_out/r02n_pytest.log
