mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r02c_pytest_multi.log; tail -5 gpurun_out/r02c_pytest_multi.log
for shape in "512 512 512" "128 128 128"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 scripts/check_slab.py --shape $shape --steps 9 --transport fused --time-steps 500 2>&1 | grep -E "SLAB_|MISMATCH|Error|error" | head -8
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; tail -5 gpurun_out/r02c_bench_n2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02c_bench_n2.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
for k in ('halo_check', 'train_gs3d_512', 'cfg4_gs3d_128', 'halo'):
    print(k, d.get(k))
PY
