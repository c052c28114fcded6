mkdir -p gpurun_out
O=gpurun_out/r02s2c
echo "== default (register monomial sums)" > ${O}_perf.txt
python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1
for v in bwd_al; do echo "== $v" >> ${O}_perf.txt; PERCNN_B200_LIB=variants/$v.so python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1
  PERCNN_B200_LIB=variants/$v.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_data_loss_gpu.py tests/test_slab_self_gpu.py tests/test_parity_full_gpu.py -x -q -m gpu 2>&1 | tail -2 >> ${O}_perf.txt; done
python scripts/perf_upscaler.py > ${O}_perf_upscaler.txt 2>&1
timeout 300 python -m pytest tests/test_upscaler_gpu.py -x -q -m gpu 2>&1 | tail -2 >> ${O}_perf_upscaler.txt
cat ${O}_perf.txt ${O}_perf_upscaler.txt
