#!/bin/bash
mkdir -p gpurun_out
for v in "0 last" "200 last" "200 normal"; do
  set -- $v
  echo "== STAT_L2_PERSIST=$1 STAT_ATT_L2=$2"
  STAT_L2_PERSIST=$1 STAT_ATT_L2=$2 timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/err.txt | tail -1 > gpurun_out/bench.json
  grep stat gpurun_out/err.txt | head -2
  python - <<PY
import json
d=json.load(open('gpurun_out/bench.json'))
r=d['roofline']
print('value %.0f e2e %.0f ms/step %.3f att us %.2f frac %.3f cold %.2f b2b %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['avg_launch_us'],r['frac'],r['isolated_cold_l2_us'],r['isolated_back_to_back_us']))
print({k:round(v['ms_per_step'],3) for k,v in d['phases_eager'].items()})
PY
done
