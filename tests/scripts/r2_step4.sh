#!/bin/bash
# round 2, step 4: full GPU test-suite (incl. tests/test_gpu_round2.py), smoke, reference arm
set -u
O=gpurun_out
L=$O/r2_step4.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 >> $L 2>&1
tail -3 $L
