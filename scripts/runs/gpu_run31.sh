mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01e_launches_cfg1_epoch.csv python scripts/profile_cfg1_epoch.py 40 > gpurun_out/r01e_ncu.log 2>&1; tail -2 gpurun_out/r01e_ncu.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r01e_launches_cfg1_epoch.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:70]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])): print(f"{len(v):5d} x {sum(v)/len(v)/1e3:9.2f} us  {k}")
PY
