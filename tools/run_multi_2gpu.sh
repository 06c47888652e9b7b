#!/bin/bash
# two-GPU tests of the sharded renderer + a two-rank bench run (gpurun --gpus 2 -- bash tools/run_multi_2gpu.sh)
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" && mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/multi2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/multi2_pytest.log
tail -4 gpurun_out/multi2_pytest.log
if [ "$1" != "tests-only" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/multi2_bench_n2.json 2> gpurun_out/multi2_bench_n2.err
tail -1 gpurun_out/multi2_bench_n2.json | cut -c1-400
fi
