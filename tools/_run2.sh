#!/bin/bash
cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r2e_gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/r2e_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err
tail -4 gpurun_out/r2e_pytest_multi.log; cut -c1-1500 gpurun_out/r2e_bench_n2.json; tail -5 gpurun_out/r2e_bench_n2.err
