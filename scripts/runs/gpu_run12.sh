timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs3d_bwd_tma -s 2 -c 1 -f -o gpurun_out/r01_tma_bwd_256 python scripts/profile_step.py --n 256 --steps 4 --bwd > gpurun_out/p4.log 2>&1
tail -2 gpurun_out/p4.log
