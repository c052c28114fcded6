mkdir -p gpurun_out
timeout 900 python scripts/sweep_tile2d.py 2>&1 | tee gpurun_out/r02m_sweep_tile2d.txt | tail -120
