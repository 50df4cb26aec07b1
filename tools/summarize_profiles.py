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


def sass_profile(md):
    """Instruction mix and stall reasons of the attention kernel from the source page of its capture."""
    path = os.path.join(SRC, 'att_sass.csv')
    if not os.path.isfile(path):
        return
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stallcols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    byop, stall, tot, nsamp = collections.Counter(), collections.Counter(), 0, 0
    for r in data:
        if len(r) < len(hdr):
            if r and r[0] == 'Kernel Name':
                break                     # the second captured launch: one is enough
            continue
        if r[ix['Source']] == 'Source':
            continue
        toks = r[ix['Source']].split()
        op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
        n = int(r[ix['Instructions Executed']] or 0)
        byop[op] += n
        tot += n
        nsamp += int(r[ix['# Samples']] or 0)
        for c in stallcols:
            stall[c] += int(r[ix[c]] or 0)
    md.append('### att_group_kernel: executed warp instructions by opcode and warp-state samples (`--page source`)\n')
    md.append('%d warp instructions in the launch; ' % tot + ', '.join('%s %.1f%%' % (o, 100.0 * n / tot)
                                                                    for o, n in byop.most_common(10)) + '.\n')
    md.append('%d samples: ' % nsamp + ', '.join('%s %d' % (c, n) for c, n in stall.most_common(8)) + '.\n')
    mufu, ublk = byop.get('MUFU', 0), byop.get('UBLKCP', 0)
    md.append('MUFU %d (5 per 4 tanh), UBLKCP (cp.async.bulk) %d, no tensor-pipe instructions (HBM-bound '
              'streaming kernel).\n' % (mufu, ublk))


def step_traffic(tag, md):
    """Per-kernel DRAM bytes of the decode steps measured in situ: one ncu pass per launch, no replay,
    caches left alone (tools/ncu_full.sh step 2 / the step_dram capture)."""
    path = os.path.join(SRC, 'step_dram.csv')
    if not os.path.isfile(path):
        return None
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    d = collections.OrderedDict()
    for r in rows[1:]:
        e = d.setdefault(int(r[ix['ID']]), {'name': re.sub(r'\(.*', '', r[ix['Kernel Name']]).split('::')[-1],
                                            'grid': r[ix['Grid Size']]})
        e[r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
    att = [i for i in d if d[i]['name'].startswith('att_group')]
    if len(att) < 12:
        return None
    s, e = att[8], att[10]
    md.append('## DRAM traffic of two decode steps in situ (`%s_step_dram.csv`)\n' % tag)
    md.append('`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct '
              '--cache-control none --clock-control none` on `tools/one_step.py`: one pass per launch, nothing replayed, '
              'caches as the preceding kernels left them.\n')
    md.append('| # | kernel | grid | us | DRAM read MB | DRAM write MB | L2 hit % |\n|---|---|---|---|---|---|---|')
    with open(os.path.join(OUT, '%s_step_dram.csv' % tag), 'w') as f:
        f.write('index,kernel,grid,duration_us,dram_read_bytes,dram_write_bytes,l2_hit_pct\n')
        for i in d:
            x = d[i]
            f.write('%d,"%s","%s",%.2f,%.0f,%.0f,%.2f\n' % (i, x['name'], x['grid'], x['gpu__time_duration.sum'] / 1e3,
                                                         x['dram__bytes_read.sum'], x['dram__bytes_write.sum'],
                                                         x['lts__t_sector_hit_rate.pct']))
            if s <= i < e:
                md.append('| %d | %s | %s | %.1f | %.2f | %.2f | %.1f |' % (
                    i, x['name'], x['grid'], x['gpu__time_duration.sum'] / 1e3, x['dram__bytes_read.sum'] / 1e6,
                    x['dram__bytes_write.sum'] / 1e6, x['lts__t_sector_hit_rate.pct']))
    md.append('')
    a = [d[i] for i in att[2:]]
    n = len(a)
    return {'launches': n, 'dram_read_bytes_per_launch': sum(x['dram__bytes_read.sum'] for x in a) / n,
            'dram_write_bytes_per_launch': sum(x['dram__bytes_write.sum'] for x in a) / n,
            'l2_hit_pct': sum(x['lts__t_sector_hit_rate.pct'] for x in a) / n,
            'source': '%s_step_dram.csv (one ncu pass per launch, --cache-control none)' % tag}


def bench_launch_list(tag, md):
    """Launch list of bench.py itself (first 1200 launches): one replay of the captured K0 + 20-step graph."""
    path = os.path.join(SRC, 'bench_launches.csv')
    if not os.path.isfile(path):
        return
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith('=='))]
    hdr = rows[0]
    ki, vi, ui, gi = (hdr.index(k) for k in ('Kernel Name', 'Metric Value', 'Metric Unit', 'Grid Size'))
    seq = []
    for row in rows[1:]:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(',', ''))
        seq.append((re.sub(r'\(.*', '', row[ki]).split('::')[-1], row[gi], v / 1000 if row[ui] == 'ns' else v))
    with open(os.path.join(OUT, '%s_bench_launches.csv' % tag), 'w') as f:
        f.write('index,kernel,grid,duration_us\n')
        for i, (n, g, v) in enumerate(seq):
            f.write('%d,"%s","%s",%.2f\n' % (i, n, g, v))
    starts = [i for i, x in enumerate(seq) if x[0].startswith('meanpool')]
    if len(starts) < 4:
        return
    a, b = starts[2], starts[3]               # a graph replay (the first two passes are the eager warm-up / capture)
    step = seq[a:b]
    tot = sum(v for _, _, v in step)
    agg = collections.OrderedDict()
    for n, g, v in step:
        e = agg.setdefault((n, g), [0, 0.0])
        e[0] += 1
        e[1] += v
    md.append('## Launch list of `bench.py --steps 2 --warmup 3` itself — `%s_bench_launches.csv`\n' % tag)
    md.append('`ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 python bench.py --steps 2 --warmup 3`: the '
              'first 1200 launches of the bench command (parameter packing, eager warm-up, then replays of the captured '
              'K0 + 20-step graph, 150 kernel nodes each).  One replay = launches %d..%d, %.0f us in total under the '
              'profiler (serialised); shares agree with the eager list above.\n' % (a, b - 1, tot))
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
    bench_launch_list(tag, md)
    att = summarize(os.path.join(SRC, 'prof_att.ncu-rep'), 'att_group_kernel — the HBM/L2-bound attention kernel (full set: replayed, caches flushed between passes = cold numbers)', md)
    insitu = step_traffic(tag, md)
    sass_profile(md)
    summarize(os.path.join(SRC, 'prof_gemm.ncu-rep'), 'gemm_tf32x3_kernel — K0 projections and per-step GEMMs', md)
    with open(os.path.join(OUT, '%s_ncu_summary.md' % tag), 'w') as f:
        f.write('\n'.join(md) + '\n')
    if att:
        a = att[-1]
        rd = to_bytes(*a['dram__bytes_read.sum'])
        wr = to_bytes(*a['dram__bytes_write.sum'])
        rec = {'kernel': a['kernel'], 'dram_bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr,
               'source': '%s_ncu_summary.md (ncu --set full, one launch, caches flushed by ncu: every byte from DRAM; '
                         'B=64 T=26 R=8 H=512)' % tag}
        if insitu:
            rec['in_situ'] = insitu
        # the commit whose att_group.cu the capture was taken from (bench.py prints it next to roofline.traffic)
        try:
            rec['captured_at_commit'] = subprocess.run(
                ['git', 'log', '-1', '--format=%h', '--',
                 'video_description_with_spatial_temporal_attention_b200/csrc/att_group.cu'],
                capture_output=True, text=True, cwd=ROOT).stdout.strip()
        except Exception:
            rec['captured_at_commit'] = '?'
        with open(os.path.join(OUT, 'att_step_traffic.json'), 'w') as f:
            json.dump(rec, f, indent=1)
    print('\n'.join(md[:60]))


if __name__ == '__main__':
    main()
