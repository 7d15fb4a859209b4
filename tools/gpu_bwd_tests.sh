#!/bin/bash
# Runs each backward test group in its own process (a trapped kernel poisons only its own group); logs under gpurun_out/.
mkdir -p gpurun_out
out=gpurun_out/bwd_tests.log
: > $out
for t in "$@"; do
  echo "=== $t" >> $out
  timeout 600 python -m pytest tests/test_gpu_backward.py -q -k "$t" --timeout 180 -p no:cacheprovider 2>&1 | tail -60 >> $out
done
tail -c 6000 $out
