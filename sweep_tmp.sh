# C2: ncu full on the production kernel at full spp? too long under replay -> spp 8 profile + launch list at full spp (no replay)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:megakernel -c 2 --csv --log-file gpurun_out/c2_full_dram.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:megakernel -s 1 -c 1 -f -o gpurun_out/prof_c4 \
    python bench.py --workload c4 --steps 1 --warmup 1 --spp 1 --no-cpu > gpurun_out/ncu_c4.log 2>&1
timeout 600 python bench.py --workload c5 --frames 300 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2_v3.csv \
    python bench.py --steps 2 --warmup 1 --spp 8 --no-cpu > gpurun_out/ncu_launch.log 2>&1
