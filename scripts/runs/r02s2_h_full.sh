mkdir -p gpurun_out
O=gpurun_out/r02s2h
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_pytest.log 2>&1; tail -4 ${O}_pytest.log
python scripts/perf_slab_train_self.py > ${O}_slab_train.txt 2>&1; cat ${O}_slab_train.txt
python scripts/perf_bwd.py > ${O}_perf_bwd.txt 2>&1; cat ${O}_perf_bwd.txt
