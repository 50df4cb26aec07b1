#!/bin/bash
# parity suite + bench line (what the round's final numbers come from)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_all_s3.log
timeout 900 python bench.py --steps 50 --warmup 5 2>gpurun_out/bench_r2_t.err | tail -1 > gpurun_out/bench_r2_t.json
tail -3 gpurun_out/bench_r2_t.err
