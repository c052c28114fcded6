timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pi_k5 -s 2 -c 2 -f -o gpurun_out/r01_k5_512 python scripts/profile_k5.py > gpurun_out/p5.log 2>&1
tail -1 gpurun_out/p5.log
