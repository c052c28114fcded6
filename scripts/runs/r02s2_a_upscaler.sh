mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_upscaler_gpu.py tests/test_stage2_library.py tests/test_modules_gpu.py "tests/test_slab_self_gpu.py::test_self_ring_initial_state_from_the_sharded_upscaler" -x -q -m gpu > gpurun_out/r02s2a_pytest_new.log 2>&1; tail -15 gpurun_out/r02s2a_pytest_new.log
timeout 300 python scripts/perf_upscaler.py > gpurun_out/r02s2a_perf_upscaler.txt 2>&1; cat gpurun_out/r02s2a_perf_upscaler.txt | tail -12
