#!/bin/bash
# One GPU session: tests, bench, ncu launch list, ncu full capture of the top kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --spp 8 --no-cpu > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:megakernel -s 1 -c 1 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
