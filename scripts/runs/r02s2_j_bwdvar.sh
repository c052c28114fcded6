mkdir -p gpurun_out
O=gpurun_out/r02s2j
echo "== default" > ${O}_perf.txt
python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1
for v in bwd_earlyh bwd_nopf bwd_earlyh_nopf; do echo "== $v" >> ${O}_perf.txt; PERCNN_B200_LIB=variants/$v.so python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1; done
echo "== default again" >> ${O}_perf.txt
python scripts/perf_bwd.py >> ${O}_perf.txt 2>&1
grep -E "==|512" ${O}_perf.txt
