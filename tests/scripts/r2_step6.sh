#!/bin/bash
# round 2, step 6: graph-GRU with fused message weights (one recurrent GEMM + one exchange per step)
set -u
O=gpurun_out
L=$O/r2_step6.log
: > $L
timeout 900 python -m pytest tests -m gpu -x -q -k "GGRU or ggru or gru" 2>&1 | tail -8 >> $L
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;128,128;64,128;64,64" "A3GC_TC_OPT=0" 1024 40 fp32 GGRU >> $L 2>&1
timeout 300 python tests/prof_sweep.py "256,512;64,128" "A3GC_TC_OPT=0" 1024 40 bf16 GGRU >> $L 2>&1
timeout 600 python bench.py --variant GGRU --seq-len 600 --no-cpu-baseline >> $L 2>&1
tail -3 $L
