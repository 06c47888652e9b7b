timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
echo "c4 host PLOC"; run --workload c4 --steps 2
echo "c4 gpu LBVH"; run --workload c4 --steps 2 --gpu-bvh
echo "c2"; run
