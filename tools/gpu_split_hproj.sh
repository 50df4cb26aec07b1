#!/bin/bash
# greedy decoding with the readout chain forked right after the gates (STAT_SPLIT_HPROJ=1): parity subset + A/B timing
mkdir -p gpurun_out
STAT_SPLIT_HPROJ=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 250 -p no:cacheprovider -k "golden_greedy or config2_greedy or caption_stream or baseline_size" 2>&1 | tail -3 | tee gpurun_out/t_split_hproj.log
for v in 0 1 0 1; do
  STAT_SPLIT_HPROJ=$v timeout 200 python tools/quick_value.py 2>&1 | tail -2 | head -1
done | tee gpurun_out/sweep_split_hproj.txt
