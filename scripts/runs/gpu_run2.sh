set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_512.csv python scripts/profile_step.py --n 512 --steps 12 > gpurun_out/p1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gs3d_fwd_tma -s 4 -c 2 -f -o gpurun_out/r01_tma_512 python scripts/profile_step.py --n 512 --steps 8 > gpurun_out/p2.log 2>&1
ncu --set full --clock-control none -k regex:k_gs3d_fwd_tma -s 4 -c 1 -f -o gpurun_out/r01_tma_128 python scripts/profile_step.py --n 128 --steps 8 > gpurun_out/p3.log 2>&1
tail -3 gpurun_out/p1.log gpurun_out/p2.log gpurun_out/p3.log
ls -la gpurun_out
