#!/bin/bash
# round 2, step 49: G-GRU kernel, MMA issue loop on the whole warp (uniform operands, elected lane issues)
set -u
O=gpurun_out
L=$O/r2_step49.log
: > $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -m gpu -x -q -k "GGRU or ggru or gru or smoke or chain" 2>&1 | tail -2 >> $L
timeout 600 python tests/prof_sweep.py "256,512;256,256;128,256;128,128;64,128;64,64" "A3GC_TC_OPT=0" 1024 40 fp32 GGRU >> $L 2>&1
timeout 600 python tests/prof_sweep.py "256,512;64,128" "A3GC_TC_OPT=0" 1024 40 bf16 GGRU >> $L 2>&1
for s in "256 512" "256 256" "64 128"; do timeout 300 python tests/prof_gru_trace.py $s 2>&1 | grep -E "G-GRU|step 4|mma|MMA thread|producer|per step" >> $L; done
timeout 600 python bench.py --variant GGRU --seq-len 600 --no-secondary --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 >> $L
tail -3 $L
