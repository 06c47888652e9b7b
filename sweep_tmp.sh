timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --workload c4 --steps 2 --warmup 3 --cpu-spp 1 2>&1 | tail -1
timeout 600 python bench.py --workload c5 --frames 100 2>&1 | tail -1
