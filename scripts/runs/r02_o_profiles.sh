mkdir -p gpurun_out
# every launch of a short default bench run with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench_n1.log 2>&1
tail -2 gpurun_out/r02_launches_bench_n1.log | cut -c1-200
# full captures of the flagship kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gs3d|k_tile2d|k_pi_k5" -o gpurun_out/r02_ncu_kernels -f python scripts/ncu_kernels.py > gpurun_out/r02_ncu_kernels.log 2>&1
tail -2 gpurun_out/r02_ncu_kernels.log
ls -la gpurun_out/r02_ncu_kernels.ncu-rep
# the real bench line (not under a profiler)
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -2 gpurun_out/r02_bench_n1.err | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
