timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"; }
for w in 22 26 30; do echo "ch wait=$w"; BVR_MK_WAIT=$w run; done
echo cta-wavefront; run --kernel cta-wavefront
echo c4; run --workload c4 --steps 2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 1 -c 1 -f -o gpurun_out/prof_v3d \
    python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu > gpurun_out/ncu_full.log 2>&1
