set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -3
