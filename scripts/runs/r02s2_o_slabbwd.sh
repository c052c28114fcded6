mkdir -p gpurun_out
O=gpurun_out/r02s2o
timeout 900 python -m pytest tests/test_slab_self_gpu.py tests/test_parity_gpu.py tests/test_data_loss_gpu.py tests/test_parity_full_gpu.py -x -q -m gpu 2>&1 | tail -2
python scripts/perf_slab_train_self.py > ${O}_slab_train.txt 2>&1; cat ${O}_slab_train.txt
python scripts/perf_bwd.py 2>&1 | tail -2
