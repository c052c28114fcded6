mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gs3d_fwd -o gpurun_out/r02f_slab_vs_periodic -f python scripts/ncu_slab_vs_periodic.py > gpurun_out/r02f_ncu.log 2>&1; tail -3 gpurun_out/r02f_ncu.log
ncu -i gpurun_out/r02f_slab_vs_periodic.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,launch__registers_per_thread,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__cycles_active.avg,sm__inst_executed_pipe_fma.sum,launch__grid_size,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_membar_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_sleeping_per_warp_active.pct,smsp__issue_active.avg.pct_of_peak_sustained_active > gpurun_out/r02f_ncu_raw.csv 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02f_ncu_raw.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID']
if hdr:
    h=rows[hdr[0]]
    for r in rows[hdr[0]+2:]:
        d=dict(zip(h,r))
        print(d.get('Kernel Name','')[:40], {k.split('__',1)[-1][:38]:v for k,v in d.items() if '__' in k})
PY
