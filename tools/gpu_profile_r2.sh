#!/bin/bash
# Round-2 evidence run (one GPU): bench lines, ncu launch lists of the inference bench and of one training step, single-op
# timings of the norm / elementwise / loss kernels, and ncu --set full captures of the dominant forward and backward kernels.
# Every ncu command runs only after its plain command has exited 0.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-eager-baseline > gpurun_out/r2_ncu_bench.log 2>&1
echo "infer launch list rc=$?"
python tools/bench_train.py --batch 8 --breakdown --eager-baseline > gpurun_out/r2_train_b8.json 2> gpurun_out/r2_train_b8.err || exit 1
NL=$(python -c "import json; print(json.load(open('gpurun_out/r2_train_b8.json'))['launches_per_step_eager'])")
bash tools/ncu_train_launches.sh 8 $NL > gpurun_out/r2_train_launch_summary_b8.txt 2>&1
echo "train launch list rc=$? ($NL launches per step)"
python tools/bench_train.py --batch 32 --breakdown > gpurun_out/r2_train_b32.json 2> gpurun_out/r2_train_b32.err
python tools/bench_ops.py band32c1 band32c2 band64c1 gnstats32 gnstats128 gnapply32 gnapply128 l1l2 kl sample cin1 cin4 cout1 cout4 upn64 upn128 attn1k > gpurun_out/r2_ops.log 2>&1
cap() {  # name kernel-regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r2prof_$name "$@" > gpurun_out/r2p_$name.log 2>&1
  echo "capture $name rc=$?"
}
cap band32c2 conv3x3_band 3 python tools/bench_ops.py band32c2
cap band32c1 conv3x3_band 3 python tools/bench_ops.py band32c1
cap gnstats32 gn_stats_kernel 3 python tools/bench_ops.py gnstats32
cap gnapply32 gn_apply_kernel 3 python tools/bench_ops.py gnapply32
cap l1l2 l1l2_partial 3 python tools/bench_ops.py l1l2
cap cin1 conv3x3_cin1 3 python tools/bench_ops.py cin1
cap cout1 conv3x3_fewcout 3 python tools/bench_ops.py cout1
cap wgrad_col wgrad3x3_col_kernel 45 python tools/bench_train.py --batch 8 --steps 1 --warmup 1 --no-graph
cap gn_bwd_reduce gn_bwd_reduce_kernel 50 python tools/bench_train.py --batch 8 --steps 1 --warmup 1 --no-graph
cap dgrad_umma conv_umma_kernel 140 python tools/bench_train.py --batch 8 --steps 1 --warmup 1 --no-graph
ls -la gpurun_out/r2prof_* | awk '{print $5, $9}'
