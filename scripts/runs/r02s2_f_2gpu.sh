mkdir -p gpurun_out
O=gpurun_out/r02s2f
timeout 1200 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > ${O}_pytest_multi.log 2>&1; tail -6 ${O}_pytest_multi.log
for k in 1 2 4; do
PERCNN_SLAB_TB_K=$k timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29620 scripts/check_slab.py --shape 128 128 128 --steps 20 --transport fused --time-steps 500 2>&1 | grep -E "SLAB_" | sed "s/^/K=$k /" >> ${O}_cfg4_2gpu.txt
done
cat ${O}_cfg4_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 5 --warmup 3 > ${O}_bench_n2.json 2> ${O}_bench_n2.err; tail -2 ${O}_bench_n2.err | cut -c1-300
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/r02s2f_bench_n2.json') if l.startswith('{')][-1])
print('value', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
for k in ('halo_check', 'train_gs3d_512', 'cfg4_gs3d_128'):
    v = d.get(k)
    if isinstance(v, dict): v = {a: b for a, b in v.items() if a not in ('note', 'includes', 'against')}
    print(k, v)
PY
