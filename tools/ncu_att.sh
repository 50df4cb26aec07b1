#!/bin/bash
# ncu --set full capture (with source) of the attention kernel inside one eager bench step
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_ -s 3 -c 1 -f -o gpurun_out/prof_att python tools/one_step.py > gpurun_out/ncu_att.log 2>&1
tail -2 gpurun_out/ncu_att.log
ncu -i gpurun_out/prof_att.ncu-rep --page source --csv --print-source sass > gpurun_out/att_sass.csv 2>/dev/null
ncu -i gpurun_out/prof_att.ncu-rep --page details > gpurun_out/att_details.txt 2>/dev/null
ls -la gpurun_out
