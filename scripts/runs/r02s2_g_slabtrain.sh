mkdir -p gpurun_out
O=gpurun_out/r02s2g
echo "== default (reg mono + aligned entry)" > ${O}_slab_train.txt
python scripts/perf_slab_train_self.py >> ${O}_slab_train.txt 2>&1
for v in bwd_old bwd_rm_only bwd_al_only; do echo "== $v" >> ${O}_slab_train.txt; PERCNN_B200_LIB=variants/$v.so python scripts/perf_slab_train_self.py >> ${O}_slab_train.txt 2>&1; done
cat ${O}_slab_train.txt
python scripts/perf_upscaler.py > ${O}_perf_upscaler.txt 2>&1; cat ${O}_perf_upscaler.txt
timeout 300 python -m pytest tests/test_upscaler_gpu.py tests/test_slab_self_gpu.py -x -q -m gpu 2>&1 | tail -2
