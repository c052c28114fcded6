for i in 1 2; do
echo "== new fwd kernel"; timeout 60 python scripts/perf_fwd.py 2>&1 | tail -2
echo "== old fwd kernel (seam wrap logic)"; PERCNN_B200_LIB=$PWD/gpurun_ab_oldfwd.so timeout 60 python scripts/perf_fwd.py 2>&1 | tail -2
done
