mkdir -p gpurun_out
O=gpurun_out/r02s2e
timeout 900 python -m pytest tests/test_slab_self_gpu.py -x -q -m gpu > ${O}_pytest.log 2>&1; tail -5 ${O}_pytest.log
timeout 600 python scripts/perf_slab_small.py > ${O}_slab_small.txt 2>&1; head -2 ${O}_slab_small.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_up_launches.csv python scripts/prof_upscaler.py > ${O}_up_prof.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02s2e_up_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
for r in rows[1:]:
    print(f"{float(r[vi].replace(',','')):12.1f}  {r[ki][:90]}")
PY
