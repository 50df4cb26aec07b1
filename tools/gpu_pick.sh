#!/bin/bash
# round-2 session 3: register-resident vocabulary reduction / fused beam bookkeeping -- parity subset, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 400 -p no:cacheprovider -k "beam" 2>&1 | tail -8 | tee gpurun_out/t_pick.log
for nt in ${NTS:-512}; do
  STAT_PICK_NT=$nt timeout 300 python tools/beam_phases.py 2>&1 | tail -14
done | tee gpurun_out/sweep_pick.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/beam_launches.csv python tools/beam_phases.py > gpurun_out/beam_ncu_stdout.log 2>&1
