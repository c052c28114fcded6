mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_final_pytest_gpu.log
python scripts/perf_bwd.py 2>&1 | tail -2
