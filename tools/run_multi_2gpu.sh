#!/bin/bash
cd $GRAFT_REPO_ROOT
nvidia-smi -L > gpurun_out/r3f_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cases.py -m gpu -x -q -rs > gpurun_out/r3f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log
tail -6 gpurun_out/r3f_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3f_bench_n2.json 2> gpurun_out/r3f_bench_n2.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/r3f_bench_n1.json 2> gpurun_out/r3f_bench_n1.err
python - <<'PY'
import json
for f in ("n1","n2"):
    try:
        j=json.loads(open(f"gpurun_out/r3f_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, j["n_gpus"], round(j["ms_per_step"],3), round(j["value"]), round(j["e2e"]["ms_per_step"],3), j.get("multi_gpu_parity"), j.get("multi_gpu_parity_detail"), {k:(round(v["ms_per_step"],3), v["parity"]) for k,v in j.get("extra",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
tail -5 gpurun_out/r3f_bench_n2.err
