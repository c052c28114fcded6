mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_slab_self_gpu.py -q -x 2>&1 | tail -3
for v in "PERCNN_FUSED_DEBUG=3" ""; do
echo "== $v"; env $v STEPS=200 timeout 300 python scripts/perf_slab_self.py 2>&1 | grep -E "^\((512|64|256), 512|^\((128|16), 128"
done 2>&1 | tee gpurun_out/r02g_variants.txt
