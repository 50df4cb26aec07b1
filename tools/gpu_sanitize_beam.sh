#!/bin/bash
# compute-sanitizer over the beam-search tests (att_clip_kernel, pick_regs_kernel, beam_select_kernel with its gather)
mkdir -p gpurun_out
SEL='beam_shared or beam_edge or beam_k1 or golden_beam'
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_beam_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 450 -p no:cacheprovider -k "$SEL" 2>&1 | tail -3 | tee gpurun_out/sanitizer_beam_memcheck.out
tail -2 gpurun_out/sanitizer_beam_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_beam_racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 450 -p no:cacheprovider -k "beam_shared and not 150" 2>&1 | tail -3 | tee gpurun_out/sanitizer_beam_racecheck.out
tail -2 gpurun_out/sanitizer_beam_racecheck.log
