#!/bin/bash
# Regenerates the round's bench artifacts on a B200 (run under gpurun from the repo root); outputs in gpurun_out/final3_*
set -u
O=gpurun_out
python bench.py > $O/final3_bench.json 2> $O/final3_bench.err
python bench.py --impl reference > $O/final3_reference.json 2>> $O/final3_bench.err
python bench.py --workload train > $O/final3_train.json 2>> $O/final3_bench.err
: > $O/final3_sidelines.jsonl
for vp in "AAGC bf16" "AGC bf16" "A3GC bf16" "AAGC fp32" "AGC fp32" "GGRU fp32"; do
  set -- $vp
  python bench.py --variant $1 --precision $2 --no-cpu-baseline >> $O/final3_sidelines.jsonl 2>> $O/final3_bench.err
done
python bench.py --variant GGRU --seq-len 600 --no-cpu-baseline >> $O/final3_sidelines.jsonl 2>> $O/final3_bench.err
: > $O/final3_traces.txt
for shape in "256 512" "256 256" "128 256" "128 128" "64 128" "64 64"; do
  A3GC_TC_TRACE=1 python tests/prof_tc.py $shape 1024 40 2>&1 | sed -n 1,3p >> $O/final3_traces.txt
  A3GC_TC_TRACE=1 python tests/prof_tc.py $shape 1024 40 2>&1 | grep "SM clock" >> $O/final3_traces.txt
done
python - <<'PY'
import json
for n in ("bench", "reference", "train"):
    try:
        d = json.load(open(f"gpurun_out/final3_{n}.json"))
        print(n, round(d["value"]), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
for line in open("gpurun_out/final3_sidelines.jsonl"):
    d = json.loads(line); print(d["config"]["workload"][:60], round(d["value"]), round(d["e2e"]["value"]))
PY
