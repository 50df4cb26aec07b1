#!/bin/bash
# ncu captures of the hot kernels inside one eager bench step (K0 + 20 greedy steps at B=64)
mkdir -p gpurun_out
# 1. full section set of the attention kernel (replayed, caches flushed between passes: cold numbers)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_group -s 3 -c 2 -f -o gpurun_out/prof_att python tools/one_step.py > gpurun_out/ncu_att.log 2>&1
tail -2 gpurun_out/ncu_att.log
# 2. DRAM bytes of every attention launch in situ: one pass, no replay, caches left alone -> what the
#    persisting-L2 carve-out really saves from step to step
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none -k regex:att_group --csv --log-file gpurun_out/att_dram.csv python tools/one_step.py > gpurun_out/ncu_att2.log 2>&1
tail -1 gpurun_out/ncu_att2.log
# 3. K0 GEMMs are launches 1..6 of gemm_tf32x3 after the 1 prepare launch; then the per-step skinny ones
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 10 -f -o gpurun_out/prof_gemm python tools/one_step.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
# 4. launch list
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/one_step.py > gpurun_out/ncu_stdout.log 2>&1
tail -1 gpurun_out/ncu_stdout.log
ncu -i gpurun_out/prof_att.ncu-rep --page source --csv --print-source sass > gpurun_out/att_sass.csv 2>/dev/null
# 5. DRAM bytes / L2 hit rate of every launch of the step in situ (one pass each, caches untouched)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none -c 400 --csv --log-file gpurun_out/step_dram.csv python tools/one_step.py > /dev/null 2>&1
