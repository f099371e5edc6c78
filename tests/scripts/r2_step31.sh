#!/bin/bash
# round 2, step 31: dzm written by the backward chain in mixed form (no split pass)
set -u
O=gpurun_out
L=$O/r2_step31.log
: > $L
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | grep -E "rel_l2|passed|failed|Error" | head -8 >> $L
for m in fused; do
  echo "== A3GC_TRAIN_ADJ=$m" >> $L
  A3GC_BWD_TRACE=1 A3GC_TRAIN_ADJ=$m timeout 600 python tests/prof_train.py 256 12 3 256 200 14 2>&1 | grep -E "iter 2|a3gc|gemm|sgemm|Kernel|elementwise|bwd trace" | sed 's/  *0.00%  *0.000us  *0.00%  *0.000us  *0.000us//' | cut -c1-150 >> $L
done
timeout 600 python tests/diag_train_parity.py 2>&1 | grep -E "MMA=4" >> $L
timeout 900 python bench.py --workload train --no-cpu-baseline 2>&1 | tail -1 >> $L
tail -3 $L
