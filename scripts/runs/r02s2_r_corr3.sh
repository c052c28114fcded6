mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_upscaler_gpu.py tests/test_modules_gpu.py "tests/test_slab_self_gpu.py::test_self_ring_initial_state_from_the_sharded_upscaler" -x -q -m gpu 2>&1 | tail -3
python scripts/perf_upscaler.py > gpurun_out/r02s2r_perf_upscaler.txt 2>&1; cat gpurun_out/r02s2r_perf_upscaler.txt
