#!/bin/bash
set -u
L=gpurun_out/r2_dbg.log
: > $L
for env in "A3GC_TC_STAGES=4" "A3GC_TC_STAGES=5" "A3GC_TC_STAGES=6" "A3GC_TC_STAGES=7" "A3GC_TC_STAGES=8"; do
  echo "== $env" >> $L
  env $env timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "net_matches" 2>&1 | grep -E "AssertionError:|passed|failed" | cut -c1-160 >> $L
done
tail -3 $L
