#!/bin/bash
# ncu --set full captures of the hot kernels (one eager bench step: K0 + 20 greedy steps)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_stream -s 3 -c 2 -f -o gpurun_out/prof_att python tools/one_step.py > gpurun_out/ncu_att.log 2>&1
tail -2 gpurun_out/ncu_att.log
# K0 GEMMs are launches 1..6 of gemm_tf32x3 after the 1 prepare launch; then the per-step skinny ones
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 10 -f -o gpurun_out/prof_gemm python tools/one_step.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/one_step.py > gpurun_out/ncu_stdout.log 2>&1
tail -1 gpurun_out/ncu_stdout.log
