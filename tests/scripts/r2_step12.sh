#!/bin/bash
# round 2, step 12: packed-weight cache + whole GPU suite + smoke + headline bench (regression check)
set -u
O=gpurun_out
L=$O/r2_step12.log
: > $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1
timeout 900 python bench.py --no-cpu-baseline > $O/r02_bench_step12.json 2>> $L
tail -c 700 $O/r02_bench_step12.json >> $L
tail -3 $L
