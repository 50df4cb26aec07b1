#!/bin/bash
# parity suite + bench + ncu launch list of one bench step
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -80 | tee gpurun_out/t_all.log
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/one_step.py > gpurun_out/ncu_stdout.log 2>&1
tail -3 gpurun_out/ncu_stdout.log
