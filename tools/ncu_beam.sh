#!/bin/bash
# full-set ncu capture of one launch each of the beam-search kernels (att_clip_kernel, pick_regs_kernel, beam_select_kernel)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name regex:'att_clip|pick_regs|beam_select' --launch-skip 30 -c 3 -f -o gpurun_out/prof_beam python tools/beam_phases.py > gpurun_out/ncu_beam.log 2>&1
tail -2 gpurun_out/ncu_beam.log
