timeout 120 python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
timeout 300 python scripts/perf_bwd.py
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bwd_256.csv python scripts/profile_step.py --n 256 --steps 4 --bwd > /dev/null 2>&1
grep -E "k_gs3d_bwd|k_monomial" gpurun_out/r01_launches_bwd_256.csv | awk -F'","' '{print $5, $NF}' | tail -6
timeout 300 python scripts/perf_cfgs.py 2>&1 | grep -E "f\+b"
