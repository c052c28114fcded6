mkdir -p gpurun_out
O=gpurun_out/r02s2d
timeout 900 python -m pytest tests/test_slab_self_gpu.py tests/test_upscaler_gpu.py -x -q -m gpu > ${O}_pytest.log 2>&1; tail -8 ${O}_pytest.log
timeout 600 python scripts/perf_slab_small.py > ${O}_slab_small.txt 2>&1; cat ${O}_slab_small.txt
python scripts/perf_upscaler.py > ${O}_perf_upscaler.txt 2>&1; cat ${O}_perf_upscaler.txt
python scripts/perf_bwd.py > ${O}_perf_bwd.txt 2>&1; tail -2 ${O}_perf_bwd.txt
