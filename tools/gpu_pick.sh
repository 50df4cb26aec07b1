#!/bin/bash
# round-2 session 3: full parity suite, bench line, launch lists with the session's kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_all_s3.log
timeout 900 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_r2_t.err | tail -1 > gpurun_out/bench_r2_t.json
tail -3 gpurun_out/bench_r2_t.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/one_step.py > gpurun_out/ncu_stdout.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --cache-control none --clock-control none -c 400 --csv --log-file gpurun_out/step_dram.csv python tools/one_step.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 600 --csv --log-file gpurun_out/beam_launches.csv python tools/beam_phases.py > gpurun_out/beam_ncu_stdout.log 2>&1
timeout 300 python tools/train_bench.py --phases 2>&1 | tail -30 > gpurun_out/train_phases_s3.txt
