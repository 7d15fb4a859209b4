#!/bin/bash
# Runs each backward test group in its own process (a trapped kernel poisons only its own group); logs under gpurun_out/.
mkdir -p gpurun_out
out=gpurun_out/bwd_tests.log
: > $out
for t in "$@"; do
  echo "=== $t" >> $out
  timeout 900 python -m pytest tests/test_gpu_backward.py -q -k "$t" --timeout 300 -p no:cacheprovider --tb=short -rP 2>&1 | grep -v "^\s*$" | tail -${TAILN:-80} >> $out
done
tail -c ${TAILC:-9000} $out
