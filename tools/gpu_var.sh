#!/bin/bash
# bench lines under environment switches: VARIANTS="A=1,B=0 A=0,B=1 ..."
mkdir -p gpurun_out
for v in $VARIANTS; do
  echo "== $v"
  env ${v//,/ } timeout 600 python bench.py --steps ${STEPS:-20} --warmup 3 2>gpurun_out/err.txt | tail -1 > gpurun_out/bench_var.json
  tail -2 gpurun_out/err.txt
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_var.json'))
r=d['roofline']
print('value %.0f e2e %.0f ms/step %.3f att us %.2f frac %.3f cold %.2f b2b %.2f'%(d['value'],d['e2e']['value'],d['ms_per_step'],r['avg_launch_us'],r['frac'],r['isolated_cold_l2_us'],r['isolated_back_to_back_us']))
PY
done
