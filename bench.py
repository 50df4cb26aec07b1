#!/usr/bin/env python
"""Benchmark of the caption-decoding hot path (BASELINE.json metric: captions/sec at
B=64, T=26, R=8, len=20).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one batch of B=64 synthetic MSVD-shaped clips through the whole path:
K0 (feature projections from the raw fp32 features) + 20 greedy decode steps
(BASELINE config 2).  `value` times the captured launch sequence with the features
already in HBM; `e2e` times the public host API with pinned-host features, H2D
copies and the D2H read of tokens/scores inside the timed region.  N > 1: one
process per GPU (torchrun), clips sharded, no collective on the data path (weak
scaling); the only communication is the max-over-ranks of the elapsed time.

--impl reference times the reference's CPU implementation of the same path: the
CPU oracle in *faithful* mode (per clip, projections recomputed on every f_next
like the compiled Theano function) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

B, T, R, MAXLEN = 64, 26, 8, 20
METRIC = 'captions/sec (B=64, T=26, R=8, len=20)'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as fh:
            v = json.load(fh).get('hbm_gbs')
        if isinstance(v, dict):                      # tolerate {"hbm_gbs": {"value": ...}}
            v = v.get('value', v.get('gbs'))
        v = float(v)
        if v > 0:
            return v, 'measured (MEASURED_PEAKS.json)'
    except Exception:
        pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def make_inputs(options, seed):
    from video_description_with_spatial_temporal_attention_b200 import synthetic
    ctxg, mg, ctxl, ml, ctxm, mm = synthetic.make_features(B, T, R, options['ctxg_dim'], options['ctxl_dim'],
                                                           options['ctxm_dim'], seed=seed)
    return ctxg, mg, ctxl, ctxm


def all_host_threads():
    """The CPU arm uses every host core: torchrun exports OMP_NUM_THREADS=1, which would leave the BLAS
    behind numpy single-threaded.  Returns the thread count in effect."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=n)
        got = [i.get('num_threads', 1) for i in threadpool_info() if i.get('user_api') == 'blas']
        return max(got) if got else n
    except Exception:
        return n


def cpu_reference_rate(options, params, feats, n_clips, repeats=1):
    """captions/s of the faithful CPU restatement (gen_sample k=1, maxlen=20 per clip)."""
    from oracle import stat_oracle as so
    all_host_threads()
    ctxg, mg, ctxl, ctxm = feats
    f_init, f_next = so.make_sampler(params, options, hoist=False)
    ml = np.ones(ctxl.shape[1:3], 'float32')
    t0 = time.perf_counter()
    n = 0
    for _ in range(repeats):
        for b in range(n_clips):
            so.gen_sample(f_init, f_next, ctxg[b % B], mg[b % B], ctxl[b % B], ml, ctxm[b % B], mg[b % B], k=1,
                          maxlen=MAXLEN)
            n += 1
    dt = time.perf_counter() - t0
    return n / dt, dt


def cpu_train_rate(options, params, n_clips=32, repeats=2):
    """CPU figure beside the secondary train_step line: forward + backward + clipping of the gradient oracle
    (oracle/grad_oracle.py: the reference's cost graph restated in torch, autograd for `tensor.grad`; fp32, all
    host cores) on a bounded sample of the config-3 workload.  Returns (tokens/s, seconds per step)."""
    import torch
    from oracle import grad_oracle as go
    from video_description_with_spatial_temporal_attention_b200 import synthetic
    all_host_threads()
    batch = synthetic.make_batch(options, B=n_clips, T=T, R=R, L=MAXLEN, seed=99, ragged=False)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        go.cost_and_grads(params, options, batch, alpha_c=0.70602, decay_c=1e-4, clip_c=10., dtype=torch.float32)
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return float(np.asarray(batch[1]).sum()) / best, best


def run_reference(args, rank):
    import video_description_with_spatial_temporal_attention_b200 as stat
    from video_description_with_spatial_temporal_attention_b200 import synthetic
    if rank != 0:
        return
    o = stat.baseline_options()
    params = synthetic.trained_like_params(o, seed=7)
    feats = make_inputs(o, 1234)
    clips_per_step = 2
    for _ in range(min(args.warmup, 2)):
        cpu_reference_rate(o, params, feats, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_rate(o, params, feats, clips_per_step)
    dt = time.perf_counter() - t0
    val = args.steps * clips_per_step / dt
    cores = all_host_threads()
    line = {'metric': METRIC, 'value': val, 'unit': 'captions/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
            'config': {'workload': 'configs[1]: greedy gen_sample(k=1) len=20, T=26 R=8 Dg=2048 Dm=Dr=4096 H=E=512 '
                                   'V=12594 (global_proj D1)', 'batch_per_gpu': B,
                       'sample': '%d clips per step of the B=64 batch' % clips_per_step},
            'cpu_baseline': {'value': val, 'unit': 'captions/s', 'cores': cores, 'kind': 'port',
                             'sample': '%d clips/step x %d steps, faithful per-clip f_next (projections recomputed '
                                       'every step), numpy fp32 BLAS on all host cores' % (clips_per_step, args.steps)},
            'e2e': {'value': val, 'unit': 'captions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def run_train_dp(o, dev, rank, world, steps=6, warmup=2, Bt=64):
    """tokens/s of the data-parallel training step (config 4): per rank B=64 clips, L=20, T=26, R=8, BASELINE widths,
    dropout on, alpha_c / decay_c / clip_c of the shipped config, adam.  Device time (CUDA events), max over
    ranks; the all-reduce share from events around the collective."""
    import torch
    import torch.distributed as dist
    from video_description_with_spatial_temporal_attention_b200 import model_attention as ma
    from video_description_with_spatial_temporal_attention_b200.train import Trainer
    params = ma.Attention().init_params(o)
    tr = Trainer(params, o, optimizer='adam', alpha_c=0.70602, decay_c=1e-4, clip_c=10., use_noise=True, seed=1234,
                 device=dev, sync_cost=False)
    g = torch.Generator(device=dev).manual_seed(4321 + rank)
    L = MAXLEN
    ctxg = torch.randn(Bt, T, o['ctxg_dim'], device=dev, generator=g) * 0.5
    ctxl = torch.randn(Bt, T, R, o['ctxl_dim'], device=dev, generator=g) * 0.5
    ctxm = torch.randn(Bt, T, o['ctxm_dim'], device=dev, generator=g) * 0.5
    mg = torch.ones(Bt, T, device=dev)
    x = torch.randint(2, o['n_words'], (L, Bt), device=dev, generator=g)
    x[L - 1] = 0
    mask = torch.ones(L, Bt, device=dev)
    batch = (x, mask, ctxg, mg, ctxl, None, ctxm, None)
    for _ in range(warmup):
        tr.f_grad_shared(*batch)
        tr.f_update(0.01)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tr.time_allreduce = True
    tr.allreduce_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = tr.engine.launch_count()
    e0.record()
    for _ in range(steps):
        cost = tr.f_grad_shared(*batch)[0]
        tr.f_update(0.01)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ms_ar = float(np.mean([a.elapsed_time(b) for a, b in tr.allreduce_events])) if tr.allreduce_events else 0.0
    t = torch.tensor([ms, ms_ar], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_ar = float(t[0]), float(t[1])
    tokens = float(mask.sum()) * world
    return {'workload': 'configs[3]: data-parallel training step (fwd + bwd + all-reduce + clip + adam), B=%d clips per '
                        'GPU, L=%d, T=%d, R=%d, BASELINE widths, dropout on' % (Bt, L, T, R),
            'tokens_per_s': tokens / (ms * 1e-3), 'ms_per_step': ms, 'ms_allreduce': ms_ar,
            'allreduce_bytes': int(tr.flat.n) * 4, 'allreduce_GBps_bus': (2.0 * (world - 1) / world) * tr.flat.n * 4 /
            (ms_ar * 1e-3) / 1e9 if ms_ar > 0 and world > 1 else None,
            'n_gpus': world, 'steps': steps, 'warmup': warmup, 'cost': float(cost),
            'launches_per_step': int(tr.engine.launch_count() - n0) // steps,
            'timing': 'CUDA events, max over ranks; collective = torch.distributed all_reduce(SUM) over NCCL on the '
                      'flat fp32 gradient buffer, not overlapped with the backward pass'}


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE any pinned host buffer is allocated
    (first touch puts the staging pages on that node).  Returns a short description for the bench line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return {'cpus': '%d-%d (%d)' % (cpus[0], cpus[-1], len(cpus)) if cpus else '', 'bound': True}
    except Exception as e:
        return {'bound': False, 'why': str(e)[:80]}


def run_ours(args, rank, world, local_rank):
    numa = bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist
    import video_description_with_spatial_temporal_attention_b200 as stat
    from video_description_with_spatial_temporal_attention_b200 import synthetic
    from video_description_with_spatial_temporal_attention_b200.engine import Engine

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU path)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    o = stat.baseline_options()
    params = synthetic.trained_like_params(o, seed=7)
    feats = make_inputs(o, 1234 + rank)
    eng = Engine(o, device=dev)
    eng.set_params(params)
    f32 = torch.float32
    host = [torch.from_numpy(a).pin_memory() for a in feats]
    devf = [h.to(dev) for h in host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: the captured K0 + 20-step launch sequence ----
    n0 = eng.launch_count()
    out = eng.greedy_captions(*devf, maxlen=MAXLEN, use_graph=True)   # warm-up + capture
    launches_per_step = (eng.launch_count() - n0) // 2               # eager warm-up + capture
    st = eng.graph_inputs(B, T, R, MAXLEN)
    static = [st['ctxg'], st['mask'], st['ctxl'], st['ctxm']]
    for _ in range(max(args.warmup, 3)):
        eng.greedy_captions(*static, maxlen=MAXLEN)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        eng.greedy_captions(*static, maxlen=MAXLEN)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tokens = out[0].cpu().numpy()

    # ---- end to end through the host API: pinned host features -> tokens on the host ----
    h2d = sum(h.numel() * 4 for h in host)
    d2h = B * MAXLEN * 8 + B * 4 + B * 4

    def host_batches(n):
        for _ in range(n):
            yield host                                  # pinned host tensors: every batch is a real H2D

    # public streaming API: H2D of batch i+1 overlaps the decode of batch i; the captions of every
    # batch are read back to the host (numpy) inside the timed region
    for _ in eng.caption_stream(host_batches(3), MAXLEN):
        pass
    barrier()
    k_e2e = max(3, min(args.steps, 20))
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    n_read = 0
    for toks, lens, scs in eng.caption_stream(host_batches(k_e2e), MAXLEN):
        n_read += toks.shape[0]
    e3.record()
    barrier()
    assert n_read == B * k_e2e
    ms_e2e = e2.elapsed_time(e3)

    # ---- per-phase device times (eager launches, CUDA events inside the library) ----
    # The GPU is first parked in a spin kernel so that the host enqueues the whole
    # launch sequence ahead of it: event intervals then measure kernels, not launch gaps.
    def eager():
        ws, d = eng.precompute(*devf)
        eng.decode_greedy(ws, d, MAXLEN)
    eager()
    torch.cuda.synchronize()
    reps = 5
    acc = {}

    def parked():
        torch.cuda._sleep(12000000)
        eager()
    for _ in range(reps):
        ph = eng.profile(parked)
        for k, (m, c) in ph.items():
            a = acc.setdefault(k, [0.0, 0])
            a[0] += m
            a[1] += c
    phases = {k: {'ms_per_step': v[0] / reps, 'launch_groups_per_step': v[1] // reps} for k, v in acc.items()
              if v[1]}

    # ---- the attention kernel alone: cold (L2 flushed before every launch) and warm ----
    ws, d = eng.precompute(*devf)
    eng.decode_greedy(ws, d, 1)                      # leaves a valid h-projection row set in ws
    flush = torch.zeros(96 * 1024 * 1024, dtype=torch.float32, device=dev)   # 384 MB > 126 MB L2
    n_att = 20
    for _ in range(3):
        eng.attention(ws, d)
    torch.cuda.synchronize()
    pairs = []
    torch.cuda._sleep(4000000)
    for _ in range(n_att):
        flush.sum()                                   # read-only flush: no dirty lines left to write back
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.attention(ws, d)
        b.record()
        pairs.append((a, b))
    torch.cuda.synchronize()
    att_cold_us = float(np.mean([a.elapsed_time(b) for a, b in pairs])) * 1e3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(2000000)
    a.record()
    for _ in range(n_att):
        eng.attention(ws, d)
    b.record()
    torch.cuda.synchronize()
    att_warm_us = a.elapsed_time(b) / n_att * 1e3
    del flush

    # ---- secondary (BASELINE configs[4] per GPU): beam search k=5, 32 clips, maxlen 30, on the device ----
    beam = None
    try:
        bB, bk, bL = 32, 5, 30
        bf = [t[:bB].contiguous() for t in devf]
        for _ in range(2):
            eng.beam_captions(*bf, k=bk, maxlen=bL, use_graph=True)
        torch.cuda.synchronize()
        n_beam = 5
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(n_beam):
            bout = eng.beam_captions(*bf, k=bk, maxlen=bL, use_graph=True)
        b1.record()
        torch.cuda.synchronize()
        ms_beam = b0.elapsed_time(b1) / n_beam
        beam = {'workload': 'configs[4] per-GPU share: beam k=5, B=32 clips (160 decode rows), maxlen=30, K0 from raw '
                            'features, bookkeeping on the device, one CUDA graph per batch',
                'captions_per_s': bB / (ms_beam * 1e-3), 'ms_per_batch': ms_beam,
                'hypotheses_returned': int(bout[3].sum().item())}
    except Exception as e:                                   # secondary: never fail the headline line
        beam = {'error': str(e)[:200]}

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---- secondary (BASELINE configs[3]): data-parallel training, B=64 clips per GPU, one NCCL all-reduce of the
    # flat gradient per step, clip after the reduce, identical adam update on every rank.  Runs at every N (N=1 is
    # the efficiency denominator), under the process group of this bench.
    train_dp = None
    if os.environ.get('STAT_BENCH_TRAIN', '1') != '0':
        try:
            eng._graphs.clear()
            eng._ws.clear()
            torch.cuda.empty_cache()
            train_dp = run_train_dp(o, dev, rank, world)
        except Exception as e:                                   # secondary: never fail the headline line
            train_dp = {'error': str(e)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    H = o['dim']
    ctx_step_bytes = 4 * H * (2 * T * R + 4 * T) * B                 # attention-context read
    att = phases.get('step_attention')
    peak, peak_src = peaks()
    roof = None
    if att:
        dur = att['ms_per_step'] / att['launch_groups_per_step'] * 1e-3
        ach = ctx_step_bytes / dur / 1e9
        traffic = None
        traffic_src = None
        tp = os.path.join(ROOT, 'profiles', 'att_step_traffic.json')
        if os.path.isfile(tp):
            with open(tp) as fh:
                tj = json.load(fh)
                ins = tj.get('in_situ')
                # in situ (one ncu pass per launch, caches untouched) when captured, else the cold capture
                traffic = (ins['dram_read_bytes_per_launch'] + ins['dram_write_bytes_per_launch']) if ins else \
                    tj.get('dram_bytes_per_launch')
                # NOT measured in this run: the committed ncu capture of this kernel (dram__bytes_read + write per
                # launch); it is re-captured whenever the kernel changes and carries the commit it was taken at
                traffic_src = 'committed ncu capture %s, kernel source at commit %s' % (
                    (ins or tj).get('source', 'profiles/att_step_traffic.json'), tj.get('captured_at_commit', '?'))
        roof = {'bound': 'hbm', 'kernel': 'att_group_kernel (4 soft-attentions of one decode step, all rows)',
                'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': traffic,
                'traffic_source': traffic_src, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': ctx_step_bytes,
                'avg_launch_us': dur * 1e6, 'timing': 'CUDA events around each launch inside the 20-step decode '
                                                      '(eager launches, host enqueued ahead of the GPU)',
                'isolated_cold_l2_us': att_cold_us, 'isolated_cold_l2_GBps': ctx_step_bytes / att_cold_us / 1e3,
                'isolated_back_to_back_us': att_warm_us,
                'isolated_back_to_back_GBps': ctx_step_bytes / att_warm_us / 1e3}
    cpu = None
    parity = None
    if world == 1:
        rate1, dt1 = cpu_reference_rate(o, params, feats, 1)            # warm-up + calibration
        n = int(max(2, min(64, 12.0 / max(dt1, 1e-3))))
        rate, dt = cpu_reference_rate(o, params, feats, n)
        cpu = {'value': rate, 'unit': 'captions/s', 'cores': all_host_threads(), 'kind': 'port',
               'sample': '%d clips of the B=64 batch (%.1f s), faithful per-clip gen_sample(k=1, maxlen=20): f_next '
                         'recomputes the feature projections every step like the compiled Theano function; numpy '
                         'fp32 BLAS on all host cores' % (n, dt)}
        # the stronger, clearly separate CPU figure BASELINE.md asks for: projections hoisted out of the step
        # and all 64 clips decoded at once (not what the reference does; the faithful figure above is)
        from oracle import stat_oracle as so
        all_host_threads()
        so.greedy_decode_batch(params, o, *feats, MAXLEN)
        t0 = time.perf_counter()
        wt, wl, wsc = so.greedy_decode_batch(params, o, *feats, MAXLEN)
        dth = time.perf_counter() - t0
        # the timed graph's own output against the oracle (the checker, outside every timed region): token ids
        # of all 64 captions and their cumulative scores.  A kernel whose captions differ has no number.
        glen, gsc = out[1].cpu().numpy(), out[2].cpu().numpy()
        same = [bool(glen[b] == wl[b] and (tokens[b, :wl[b]] == wt[b, :wl[b]]).all()) for b in range(B)]
        serr = float(max([abs(float(gsc[b]) - float(wsc[b])) for b in range(B) if same[b]] or [0.0]))
        parity = {'checker': 'oracle/stat_oracle.greedy_decode_batch (fp32, hoisted), same features and parameters',
                  'captions_identical': int(sum(same)), 'captions': B, 'max_abs_score_err': serr,
                  'tokens_checksum_oracle': int(np.where(wt >= 0, wt, -1).astype(np.int64).sum())}
        # an fp32-vs-fp32 argmax near-tie may legitimately flip a caption; anything more is a wrong kernel
        assert sum(same) >= B - 2 and serr < 1e-3, ('greedy captions differ from the oracle', parity)
        cpu['hoisted_batched'] = {'value': B / dth, 'unit': 'captions/s',
                                  'sample': 'one B=64 batch (%.1f s): projections computed once per batch, all clips '
                                            'stepped together; numpy fp32 BLAS on all host cores' % dth}
    if cpu is not None:
        try:
            rate_t, dt_t = cpu_train_rate(o, params)
            cpu['train_step'] = {'value': rate_t, 'unit': 'tokens/s',
                                 'sample': '32 clips x 20 steps of the config-3 workload (%.1f s per step): forward + '
                                           'backward + clipping of the gradient oracle, torch fp32 on all host cores'
                                           % dt_t}
        except Exception as e:
            cpu['train_step'] = {'error': str(e)[:200]}
    # secondary (BASELINE config 3): one training step -- forward, backward, clip, adam -- at B=128, timed by
    # tools/train_bench.py in its own process so that nothing it does can cost the headline line
    train = None
    if world == 1 and os.environ.get('STAT_BENCH_TRAIN', '1') != '0':
        def train_line(extra_env):
            try:
                env = dict(os.environ)
                env.update(extra_env)
                r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'train_bench.py'), '--steps', '3',
                                    '--warmup', '2', '--phases'], capture_output=True, text=True, timeout=90, env=env)
                rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith('{')]
                if not rows:
                    return {'error': (r.stderr or 'no output')[-300:]}
                out = rows[0]
                for extra in rows[1:]:
                    out.update(extra)
                return out
            except Exception as e:
                return {'error': str(e)[:200]}
        # (one retry, kept from the hunt for the GEMM splitter race that made about 1 run in 25 of this subprocess end
        # in a device fault -- fixed; the attempt count is part of the record)
        train = train_line({})
        if 'error' in train:
            first = train['error']
            train = train_line({})
            train['attempts'] = 2
            train['first_attempt_error'] = first[-160:]
    total_bytes = B * 4 * T * (o['ctxg_dim'] + o['ctxm_dim'] + R * o['ctxl_dim']) + MAXLEN * (
        ctx_step_bytes + 41571528) + 8 * B * MAXLEN
    line = {'metric': METRIC, 'value': world * B * args.steps / (ms * 1e-3), 'unit': 'captions/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'configs[1]: single-GPU persistent-LSTM greedy decode, B=64/GPU, T=26 R=8 '
                                   'Dg=2048 Dm=Dr=4096 H=E=512 V=12594 len=20 (global_proj D1), K0 from raw '
                                   'features every batch',
                       'batch_per_gpu': B, 'parallelism': 'clips sharded, %d rank(s), no data-path collective' % world,
                       'l2': 'inputs 259 MB/step > 126 MB L2, no flush', 'params': 'random, trained-like scale',
                       'precision': 'fp32 in/out, 3xTF32 tensor-core GEMMs, fp32 attention',
                       'l2_carve_out': 'context blocks copied evict_last, everything else evict_first; persisting-L2 '
                                       'carve-out %.0f MB' % (eng.l2_persist_bytes / 1048576.0)},
            'e2e': {'value': world * B * k_e2e / (ms_e2e * 1e-3), 'unit': 'captions/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'steps': k_e2e, 'ms_per_step': ms_e2e / k_e2e,
                    # the end-to-end path is bound by the host -> device copy of the raw fp32 features
                    'h2d_GBps_per_rank': h2d / (ms_e2e / k_e2e * 1e-3) / 1e9,
                    'h2d_GBps_all_ranks': world * h2d / (ms_e2e / k_e2e * 1e-3) / 1e9,
                    'host_binding_rank0': numa},
            'gpu_launches': launches_per_step * args.steps, 'launches_per_step': launches_per_step,
            'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu, 'phases_eager': phases,
            'whole_path': {'algorithmic_bytes_per_step': total_bytes,
                           'achieved_GBps': total_bytes / (ms / args.steps * 1e-3) / 1e9,
                           'frac_of_hbm_peak': total_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
            'beam5': beam, 'train_step': train, 'l2_persist_bytes': eng.l2_persist_bytes,
            'tokens_checksum': int(tokens.astype(np.int64).sum()), 'parity': parity, 'train_dp': train_dp}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
