#!/bin/bash
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" && mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_gpus.txt
CUDA_VISIBLE_DEVICES=0,1 timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs > gpurun_out/scale_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/scale_pytest.log
tail -4 gpurun_out/scale_pytest.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_bench_n$n.json 2> gpurun_out/scale_bench_n$n.err
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/scale_bench_n1.json 2> gpurun_out/scale_bench_n1.err
python - <<'PY'
import json
for f in ("n1","n2","n4","n8"):
    try:
        j=json.loads(open(f"gpurun_out/scale_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, j["n_gpus"], round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["ms_per_step"],3), j.get("multi_gpu_parity"), j.get("multi_gpu_parity_detail"), {k:(round(v["ms_per_step"],3), v["parity"]) for k,v in j.get("extra",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/scale_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c3 --steps 2 --warmup 3 --no-extra > gpurun_out/scale_bench_c3_n8.json 2> gpurun_out/scale_bench_c3_n8.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/scale_bench_c3_n8.json").read().strip().splitlines()[-1])
print("c3_n8", round(j["ms_per_step"],3), round(j["value"]), j.get("multi_gpu_parity"), j.get("multi_gpu_parity_detail"))
PY
