# round-1 re-entry run: full GPU test suite (incl. fused data loss), smoke, bench (both arms), adjoint ncu details
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/r01b_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r01b_pytest.log
tail -5 gpurun_out/r01b_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01b_smoke.log 2>&1; tail -2 gpurun_out/r01b_smoke.log
timeout 600 python bench.py > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.err; tail -c 1500 gpurun_out/r01b_bench_n1.json; tail -3 gpurun_out/r01b_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01b_bench_ref.json 2>&1; tail -c 600 gpurun_out/r01b_bench_ref.json
timeout 200 python scripts/perf_bwd.py > gpurun_out/r01b_perf_bwd.txt 2>&1; cat gpurun_out/r01b_perf_bwd.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gs3d_bwd_tma -s 2 -c 1 -f -o gpurun_out/r01b_ncu_bwd_512 python scripts/profile_step.py --n 512 --steps 4 --bwd > gpurun_out/r01b_ncu_bwd.log 2>&1; tail -2 gpurun_out/r01b_ncu_bwd.log
ls -la gpurun_out | tail -12
