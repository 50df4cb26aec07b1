#!/bin/bash
# round-2 session 3: attention g/m prefetch experiment
mkdir -p gpurun_out
for v in - STAT_ATT_PF=1 - STAT_ATT_PF=1 STAT_ATT_CS=1 STAT_ATT_CS=1,STAT_ATT_PF=1; do
  if [ "$v" = "-" ]; then v="X_=0"; fi
  env ${v//,/ } timeout 300 python tools/quick_value.py 2>&1 | tail -2
  env ${v//,/ } timeout 300 python tools/att_time.py 2>&1 | tail -2
done | tee gpurun_out/sweep_pf.txt
