#!/bin/bash
# launch list of the bench command + one full capture of the dominant kernel (run only after the plain runs exit 0)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/prof_one.py 32 32 256 1 > gpurun_out/p1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tma -s 2 -c 1 -o gpurun_out/prof_tma32_conv2 \
    python tools/prof_one.py 32 32 256 1 >> gpurun_out/p1.log 2>&1
echo "full capture rc=$?"
