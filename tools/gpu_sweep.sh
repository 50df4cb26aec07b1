#!/bin/bash
# quick_value.py under environment switches: VARIANTS="A=1,B=0 A=0 ..." ("-" = no switch)
mkdir -p gpurun_out
for v in $VARIANTS; do
  if [ "$v" = "-" ]; then v=""; fi
  env ${v//,/ } timeout 300 python tools/quick_value.py 2>&1 | tail -2
done | tee gpurun_out/sweep.txt
