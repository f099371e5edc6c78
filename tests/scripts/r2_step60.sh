#!/bin/bash
# round 2, step 60: L2 state exchange on for the training forward too: full GPU tests + the cfg-5 line
set -u
O=gpurun_out
L=$O/r2_step60.log
: > $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
timeout 900 python bench.py --workload train --no-cpu-baseline 2>&1 | tail -1 | cut -c1-700 >> $L
tail -2 $L | cut -c1-300
