#!/bin/bash
# First GPU call for the training step: parity (default and STAT_BW_FAST variants), timing with the phase table in
# both modes.  usage: gpurun --timeout 600 -- 'bash tools/gpu_train_check.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -q --tb=short -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_train.log
timeout 200 python tools/train_bench.py --steps 5 --warmup 2 --phases > gpurun_out/train_default.json 2> gpurun_out/train_default.err
STAT_BW_FAST=1 timeout 200 python tools/train_bench.py --steps 5 --warmup 2 --phases > gpurun_out/train_fast.json 2> gpurun_out/train_fast.err
python - <<'PY'
import json
for name in ('default', 'fast'):
    try:
        rows = [json.loads(l) for l in open('gpurun_out/train_%s.json' % name) if l.startswith('{')]
        print(name, 'ms/step %.2f  tokens/s %.0f' % (rows[0]['ms_per_step'], rows[0]['value']))
        if len(rows) > 1:
            ph = rows[1]['backward_phases']
            for k, v in sorted(ph.items(), key=lambda kv: -kv[1]['ms'] if isinstance(kv[1], dict) else 0):
                print('   %-24s %8.3f ms  x%d' % (k, v['ms'], v['n']))
    except Exception as e:
        print(name, 'failed:', e)
PY
