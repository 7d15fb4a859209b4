#!/bin/bash
# Round-2 evidence run, second part (one GPU): the wide-layer kernels after the two-SM kernel, the packed-half2 prologue and
# the chained launches went in.  Every ncu command runs only after its plain command has exited 0.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-eager-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "infer launch list rc=$?"
python tools/bench_train.py --batch 8 > gpurun_out/r2_train_b8_after.json 2> gpurun_out/r2_train_b8_after.err; tail -1 gpurun_out/r2_train_b8_after.err
python tools/bench_ops.py band32c1 band32c2 band64c1 s3264 s64c1 s64c2 s12864 s128c1 p128c1 s128c2 p128c2 s128c1s s128c2s p128c2s \
    s128256 p128256 s256128 p256128 s256c1 p256c1 s256c2 p256c2 gnstats32 gnstats128 gnapply32 gnapply128 l1l2 kl sample cin1 cin4 cout1 cout4 \
    upn64 upn128 attn1k attn4k256 > gpurun_out/r2_ops_wide.log 2>&1
python tools/prof_fused.py trace 3 128-128-64-1,128-128-64-0,64-64-128-1 h16 2>&1 | cut -c1-170 > gpurun_out/r2_timeline_wide.log
python tools/prof_band.py 32 0 4 trace 2>&1 | cut -c1-200 > gpurun_out/r2_timeline_band.log
./tools/micro/mufu_rate > gpurun_out/r2_sfu_rates.log 2>&1
cap() {  # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r2prof_$name "$@" > gpurun_out/r2p_$name.log 2>&1
  echo "capture $name rc=$?"
}
cap band32c2 conv3x3_band 3 python tools/bench_ops.py band32c2
cap band32c1 conv3x3_band 3 python tools/bench_ops.py band32c1
cap s128c2 conv3x3_tma2 3 env PTIVAE_PAIR=0 python tools/bench_ops.py s128c2
cap p128c2 conv3x3_pair 3 python tools/bench_ops.py p128c2
cap p128c1 conv3x3_pair 3 python tools/bench_ops.py p128c1
cap p256c1 conv3x3_pair 3 python tools/bench_ops.py p256c1
cap s64c2 conv3x3_tma2 3 python tools/bench_ops.py s64c2
ls -la gpurun_out/r2prof_* | awk '{print $5, $9}'
