#!/bin/bash
# ncu --set full captures of the hot kernels (one eager bench step)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:att_stream -s 3 -c 2 -f -o gpurun_out/prof_att python tools/one_step.py > gpurun_out/ncu_att.log 2>&1
tail -2 gpurun_out/ncu_att.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 30 -c 5 -f -o gpurun_out/prof_gemm python tools/one_step.py > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
