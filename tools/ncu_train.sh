#!/bin/bash
# launch list of one training step (config 3) -> gpurun_out/train_launches.csv ; then a per-kernel table
# usage (GPU box): gpurun --timeout 600 -- 'bash tools/ncu_train.sh'
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/train_launches.csv python tools/train_bench.py --steps 1 --warmup 0 > gpurun_out/train_ncu_stdout.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/train_launches.csv') if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    a = agg.setdefault(r[ki][:70], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print('%-70s %5d %10.1f us %5.1f%%' % (k, n, v / 1e3, 100 * v / tot))
PY
