#!/bin/bash
# First-contact GPU run: the dense primitive alone, then the parity suite with the
# SIMT GEMM (isolates attention/recurrence bugs), then with the tcgen05 GEMM, then a
# short bench.  Every stage is bounded by its own timeout.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -k test_gemm --timeout 200 -p no:cacheprovider 2>&1 | tail -60 | tee gpurun_out/t_gemm.log
STAT_GEMM_IMPL=1 timeout 1200 python -m pytest tests -m gpu -q -k "not test_gemm" --timeout 400 -p no:cacheprovider 2>&1 | tail -80 | tee gpurun_out/t_simt.log
timeout 1200 python -m pytest tests -m gpu -q -k "not test_gemm" --timeout 400 -p no:cacheprovider 2>&1 | tail -80 | tee gpurun_out/t_tc.log
timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -30 | tee gpurun_out/bench.log
