echo "== restored fwd kernel"; timeout 60 python scripts/perf_fwd.py 2>&1 | tail -2
for i in 1 2; do
echo "== bwd current"; timeout 60 python scripts/perf_bwd.py 2>&1 | tail -2
echo "== bwd with wrap logic"; PERCNN_B200_LIB=$PWD/gpurun_ab_bwdwrap.so timeout 60 python scripts/perf_bwd.py 2>&1 | tail -2
done
