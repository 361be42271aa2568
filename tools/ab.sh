#!/bin/bash
# Same-box A/B of two builds of libconzic.so (box-to-box variance on the pool is ~10-15 %, so compare on ONE box):
#   here:      build variant A, cp conzic_b200/libconzic.so conzic_b200/lib_A.so.bin ; same for B
#   under gpurun:  bash tools/ab.sh A B "python tools/profile_step.py --ii 0 --steps 5 --breakdown"
A=$1; B=$2; CMD=$3
for rep in 1 2; do
  for v in $A $B; do
    cp conzic_b200/lib_$v.so.bin conzic_b200/libconzic.so
    echo -n "$v "; timeout 300 $CMD | cut -c1-260
  done
done
