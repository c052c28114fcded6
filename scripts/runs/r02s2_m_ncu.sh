mkdir -p gpurun_out
# every launch of a short default bench run with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/r02_launches_bench_n1.log 2>&1
tail -1 gpurun_out/r02_launches_bench_n1.log | cut -c1-160
# full captures of the flagship kernels: the report stays on the box (it is ~90 MB), the text pages come back
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_gs3d|k_tile2d|k_pi_k5|k_multi_step_slab_tb" -o /tmp/r02_ncu_kernels -f python scripts/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1
tail -1 gpurun_out/r02_ncu_kernels.log
ncu -i /tmp/r02_ncu_kernels.ncu-rep --page details > gpurun_out/r02_ncu_details_all.txt 2>/dev/null
ncu -i /tmp/r02_ncu_kernels.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_all.csv 2>/dev/null
ls -la gpurun_out/r02_ncu_details_all.txt gpurun_out/r02_ncu_raw_all.csv
