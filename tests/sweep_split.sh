#!/bin/bash
# per-shape sweep of the x-part split of the tensor-core layer kernel (A3GC_TC_SPLIT="n1%,n2%")
for shape in "256 512" "256 256" "128 256" "128 128" "64 128" "64 64"; do
  for sp in "45,35" "55,45" "50,30" "60,40" "40,40" "35,45" "30,30" "70,30" "100,0" "60,20" "50,50" "40,60"; do
    echo -n "split $sp: "; A3GC_TC_SPLIT=$sp python tests/prof_tc.py $shape 1024 40 | head -1
  done
done
