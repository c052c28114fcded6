mkdir -p gpurun_out
for v in "PERCNN_FUSED_DEBUG=3 PERCNN_SLAB_DIR=1" "PERCNN_FUSED_DEBUG=3 PERCNN_SLAB_DIR=2" "PERCNN_FUSED_DEBUG=3"; do
echo "== $v"; env $v STEPS=200 timeout 300 python scripts/perf_slab_self.py 2>&1 | grep -E "^\((512|64), 512|^\(128, 128" 
done 2>&1 | tee gpurun_out/r02e_slabdir.txt
