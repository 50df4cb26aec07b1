#!/bin/bash
# parity suite, attention trace, bench lines under a few switches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 400 -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/t_all.log
timeout 300 python tools/att_trace.py > gpurun_out/att_trace.txt 2>&1
grep -E "^run|median|max " gpurun_out/att_trace.txt
for v in ${VARIANTS:-1_1 0_1}; do
  set -- ${v//_/ }
  echo "== STAT_PDL=$1 STAT_OVERLAP=$2"
  STAT_PDL=$1 STAT_OVERLAP=$2 timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/err.txt | tail -1 > gpurun_out/bench_$1_$2.json
  tail -3 gpurun_out/err.txt
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_$1_$2.json'))
r=d['roofline']
print('value %.0f e2e %.0f ms/step %.3f att us %.2f frac %.3f cold %.2f b2b %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['avg_launch_us'],r['frac'],r['isolated_cold_l2_us'],r['isolated_back_to_back_us']))
print({k:round(v['ms_per_step'],3) for k,v in d['phases_eager'].items()})
PY
done
