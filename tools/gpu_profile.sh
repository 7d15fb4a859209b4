#!/bin/bash
# launch list of the bench command + full captures of the dominant kernels (run only after the plain runs exit 0)
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
cap() {  # name kernel-regex bench_ops-case
  python tools/bench_ops.py $3 > gpurun_out/p_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -o gpurun_out/prof_$1 \
      python tools/bench_ops.py $3 >> gpurun_out/p_$1.log 2>&1
  echo "capture $1 rc=$?"
}
cap tma32_conv1 conv3x3_tma_kernel f32c1
cap tma32_conv2 conv3x3_tma_kernel f32c2
cap tma2_128_conv2 conv3x3_tma2_kernel f128c2
cap up64 up2x_conv3x3_kernel upn64
cap tma2_sc32 conv3x3_tma2_kernel f32c2sc
cap attn attn_fwd_kernel attn1k
