timeout 600 python bench.py --workload c5 --frames 300 2>&1 | tail -1
timeout 600 python bench.py --workload c5 --frames 300 --gpu-bvh 2>&1 | tail -1
