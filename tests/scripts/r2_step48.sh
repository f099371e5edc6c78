#!/bin/bash
# round 2, step 48: bf16 mode back on one ring (the two-ring form cost AAGC bf16 9 %)
set -u
O=gpurun_out
L=$O/r2_step48.log
: > $L
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
for vp in "AAGC bf16" "AGC bf16" "A3GC bf16" "A3GC fp32"; do
  set -- $vp
  timeout 600 python bench.py --variant $1 --precision $2 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['metric'], j['dtype'], round(j['value']), round(j['roofline']['frac'],3), [round(x['ms'],1) for x in j['roofline']['launches']])" >> $L
done
tail -5 $L
