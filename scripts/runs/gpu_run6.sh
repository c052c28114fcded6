set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_512.csv python scripts/profile_step.py --n 512 --steps 12 > gpurun_out/p1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs3d_fwd_tma -s 4 -c 2 -f -o gpurun_out/r01_tma_512_v2 python scripts/profile_step.py --n 512 --steps 8 > gpurun_out/p2.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_r01_n1.json
cat gpurun_out/bench_r01_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r01_ref.json
cat gpurun_out/bench_r01_ref.json
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket"
