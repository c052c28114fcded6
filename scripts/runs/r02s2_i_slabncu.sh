mkdir -p gpurun_out
O=gpurun_out/r02s2i
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_slabtrain_launches.csv python scripts/ncu_kernels.py slabtrain > ${O}_slabtrain.log 2>&1; tail -3 ${O}_slabtrain.log
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02s2i_slabtrain_launches.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
for r in rows[1:]:
    print(f"{float(r[vi].replace(',',''))/1e3:10.1f} us  {r[ki][:100]}")
PY
