#!/bin/bash
# Runs each GPU test group in its own process (a trapped kernel poisons the CUDA context) with a timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for t in test_pack_weights test_conv_umma test_conv_umma_residual_and_stats test_conv_umma_stats_strided_modes test_conv3x3_fused test_conv3x3_fused_tma test_conv3x3_fused_tma2 test_conv3x3_fused_shortcut test_up2x_conv3x3 test_groupnorm test_conv_small_cin \
         test_conv_small_cout test_conv1x1_small_and_sigma test_attention test_attention_peaked test_attention_reads_fused_qkv_slices test_latent_sample test_losses; do
  echo "=== $t" | tee -a gpurun_out/kernels.log
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "$t and not ${t}_" -x --no-header -p no:cacheprovider 2>&1 | tail -25 >> gpurun_out/kernels.log
  echo "exit=$?" >> gpurun_out/kernels.log
done
echo "=== e2e" | tee -a gpurun_out/e2e.log
timeout 900 python -m pytest tests/test_gpu_e2e.py -q -m gpu --no-header -p no:cacheprovider -s 2>&1 | tail -120 >> gpurun_out/e2e.log
grep -E "^===|passed|failed|error|exit=" gpurun_out/kernels.log gpurun_out/e2e.log | head -80
