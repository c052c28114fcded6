mkdir -p gpurun_out
O=gpurun_out/r02s2b
timeout 900 python -m pytest tests/test_upscaler_gpu.py tests/test_parity_gpu.py tests/test_parity_full_gpu.py tests/test_data_loss_gpu.py -x -q -m gpu > ${O}_pytest.log 2>&1; tail -5 ${O}_pytest.log
python scripts/perf_upscaler.py > ${O}_perf_upscaler.txt 2>&1; tail -6 ${O}_perf_upscaler.txt
echo "== default lib" > ${O}_perf.txt
PERF_ONLY=cfg3i python scripts/perf_cfgs.py >> ${O}_perf.txt 2>&1
python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1
for v in k5mb5; do echo "== $v" >> ${O}_perf.txt; PERCNN_B200_LIB=variants/$v.so PERF_ONLY=cfg3i python scripts/perf_cfgs.py >> ${O}_perf.txt 2>&1; done
for v in bwd_rm bwd_rm_ss; do echo "== $v" >> ${O}_perf.txt; PERCNN_B200_LIB=variants/$v.so python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1
  PERCNN_B200_LIB=variants/$v.so timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_data_loss_gpu.py -x -q -m gpu -k "gs3d or tma or persistent" 2>&1 | tail -2 >> ${O}_perf.txt; done
cat ${O}_perf.txt
