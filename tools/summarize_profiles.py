"""Turn the scratch ncu captures in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_profiles.py r1      # -> profiles/r1_*.md / .csv / att_step_traffic.json
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
SRC = os.path.join(ROOT, 'gpurun_out')
KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__cluster_size', 'cluster'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'XU (MUFU) pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %'),
    ('sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active', 'tensor (tc) pipe %'),
    ('sm__inst_issued.avg.per_cycle_active', 'IPC (issued)'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]


def raw_rows(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def fmt(v, u):
    return ('%s %s' % (v, u)).strip()


def summarize(rep, title, md):
    hdr, units, rows = raw_rows(rep)
    ix = {k: i for i, k in enumerate(hdr)}
    md.append('## %s (`%s`)\n' % (title, os.path.basename(rep)))
    res = []
    for r in rows:
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']]).replace('stat::<unnamed>::', '')
        md.append('### %s  grid %s\n' % (name, r[ix['launch__grid_size']] if 'launch__grid_size' in ix else '?'))
        md.append('| metric | value |\n|---|---|')
        rec = {'kernel': name}
        for k, label in KEYS:
            if k in ix and r[ix[k]] != '':
                md.append('| %s (`%s`) | %s |' % (label, k, fmt(r[ix[k]], units[ix[k]])))
                rec[k] = (r[ix[k]], units[ix[k]])
        md.append('')
        res.append(rec)
    return res


def to_bytes(v, u):
    x = float(v.replace(',', ''))
    return x * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)


def launch_list(tag, md):
    path = os.path.join(SRC, 'launches.csv')
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui, gi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size'))
    seq = []
    for row in r:
        if len(row) <= vi:
            continue
        name = re.sub(r'\(.*', '', row[ki]).split('::')[-1]
        v = float(row[vi].replace(',', ''))
        v = v / 1000 if row[ui] == 'ns' else v
        seq.append((name, v, row[gi]))
    with open(os.path.join(OUT, '%s_launches.csv' % tag), 'w') as f:
        f.write('index,kernel,grid,duration_us\n')
        for i, (n, v, g) in enumerate(seq):
            f.write('%d,"%s","%s",%.2f\n' % (i, n, g, v))
    agg = collections.OrderedDict()
    # skip the one-off parameter packing (everything before the first meanpool launch)
    start = next((i for i, x in enumerate(seq) if x[0].startswith('meanpool')), 0)
    step = [x for x in seq[start:] if not x[0].startswith(('operator', 'vectorized', 'elementwise', 'reduce'))]
    tot = sum(v for _, v, _ in step)
    for n, v, g in step:
        a = agg.setdefault((n, g), [0, 0.0])
        a[0] += 1
        a[1] += v
    md.append('## Launch list of one eager bench step (K0 + 20 greedy steps, B=64) — `%s_launches.csv`\n' % tag)
    md.append('`ncu --metrics gpu__time_duration.sum --clock-control none` (cold caches, serialised: compare '
              'shares, not absolutes). %d launches, %.0f us in total.\n' % (len(step), tot))
    md.append('| kernel | grid | launches | total us | avg us | share |\n|---|---|---|---|---|---|')
    for (n, g), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        md.append('| %s | %s | %d | %.1f | %.2f | %.1f%% |' % (n, g, c, t, t / c, 100 * t / tot))
    md.append('')


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    os.makedirs(OUT, exist_ok=True)
    md = ['# ncu summaries, round %s\n' % tag,
          'Captured on a B200 with `tools/ncu_full.sh` (`ncu --set full --clock-control none --import-source on`) '
          'on `tools/one_step.py` (one eager pass of the bench workload).  Numbers under a profiler are evidence '
          'of *where* time and bytes go; throughput claims come from `bench.py`.\n']
    launch_list(tag, md)
    att = summarize(os.path.join(SRC, 'prof_att.ncu-rep'), 'att_stream_kernel — the HBM-bound attention kernel', md)
    summarize(os.path.join(SRC, 'prof_gemm.ncu-rep'), 'gemm_tf32x3_kernel — K0 projections and per-step GEMMs', md)
    with open(os.path.join(OUT, '%s_ncu_summary.md' % tag), 'w') as f:
        f.write('\n'.join(md) + '\n')
    if att:
        a = att[-1]
        rd = to_bytes(*a['dram__bytes_read.sum'])
        wr = to_bytes(*a['dram__bytes_write.sum'])
        with open(os.path.join(OUT, 'att_step_traffic.json'), 'w') as f:
            json.dump({'kernel': a['kernel'], 'dram_bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr,
                       'source': '%s_ncu_summary.md (ncu --set full, one launch, B=64 T=26 R=8 H=512)' % tag}, f,
                      indent=1)
    print('\n'.join(md[:60]))


if __name__ == '__main__':
    main()
