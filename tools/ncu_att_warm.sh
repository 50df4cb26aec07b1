#!/bin/bash
# full-set capture of the attention kernel with the caches left alone (replays run L2-warm): where do the
# cycles go when the copies are not the limit?
mkdir -p gpurun_out
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k regex:att_group -s 5 -c 1 -f -o gpurun_out/prof_att_warm python tools/one_step.py > gpurun_out/ncu_att_warm.log 2>&1
tail -1 gpurun_out/ncu_att_warm.log
ncu -i gpurun_out/prof_att_warm.ncu-rep --page source --csv --print-source sass > gpurun_out/att_sass_warm.csv 2>/dev/null
ncu -i gpurun_out/prof_att_warm.ncu-rep --page details > gpurun_out/att_details_warm.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|L2 Hit|Issue Slots Busy|Issued Ipc|No Eligible|Eligible Warps|Warp Cycles Per Issued" gpurun_out/att_details_warm.txt
