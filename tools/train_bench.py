"""Timing of one training step (BASELINE config 3: B=128 clips, L=20, BASELINE widths; fwd + bwd + clip + adam)
with CUDA events, per phase.  Synthetic features drawn on the device, random-init parameters.
    python tools/train_bench.py [--B 128] [--steps 2] [--warmup 1] [--phases]  ->  one JSON line (+ the phase table)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_bench.py   (config 4: B per rank)"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import video_description_with_spatial_temporal_attention_b200 as stat
from video_description_with_spatial_temporal_attention_b200 import model_attention as ma, optim
from video_description_with_spatial_temporal_attention_b200.train import Trainer


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--B', type=int, default=128)
    ap.add_argument('--L', type=int, default=20)
    ap.add_argument('--T', type=int, default=26)
    ap.add_argument('--R', type=int, default=8)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=1)
    ap.add_argument('--phases', action='store_true', help='one extra step with the per-phase timing of stat_grad_shared')
    a = ap.parse_args()
    # data parallel (BASELINE config 4) under torchrun: one rank per GPU, B clips per rank, one SUM all-reduce of the
    # flat gradient per step inside Trainer.f_grad_shared
    world, rank = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    o = stat.baseline_options()
    params = ma.Attention().init_params(o)
    tr = Trainer(params, o, optimizer='adam', alpha_c=0.70602, decay_c=1e-4, clip_c=10., use_noise=True, seed=1234 + rank)
    dev = tr.engine.device
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    B, L, T, R = a.B, a.L, a.T, a.R
    ctxg = torch.randn(B, T, o['ctxg_dim'], device=dev, generator=g) * 0.5
    ctxl = torch.randn(B, T, R, o['ctxl_dim'], device=dev, generator=g) * 0.5
    ctxm = torch.randn(B, T, o['ctxm_dim'], device=dev, generator=g) * 0.5
    mg = torch.ones(B, T, device=dev)
    x = torch.randint(2, o['n_words'], (L, B), device=dev, generator=g)
    x[L - 1] = 0
    mask = torch.ones(L, B, device=dev)
    batch = (x, mask, ctxg, mg, ctxl, None, ctxm, None)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    times = []
    lc0 = tr.engine.launch_count()
    for it in range(a.warmup + a.steps):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        cost = tr.f_grad_shared(*batch)[0]
        e1.record()
        tr.f_update(0.01)
        e2.record()
        torch.cuda.synchronize()
        if it >= a.warmup:
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        if it == 0:
            lc1 = tr.engine.launch_count()
    ms = float(np.mean([t[0] + t[1] for t in times]))
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)              # device time, max over ranks
        ms = float(tmax[0])
        if rank != 0:
            dist.destroy_process_group()
            return
    tokens = float(mask.sum()) * world
    print(json.dumps({'metric': 'training tokens/sec (config 3: fwd+bwd+clip+adam, dropout on)', 'value': tokens / ms * 1e3,
                      'unit': 'tokens/s', 'ms_per_step': ms, 'ms_grad_shared': float(np.mean([t[0] for t in times])),
                      'ms_update': float(np.mean([t[1] for t in times])), 'cost': cost, 'n_gpus': world, 'B': B, 'L': L, 'T': T, 'R': R,
                      'launches_per_step': int(lc1 - lc0), 'steps': a.steps, 'warmup': a.warmup,
                      'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}), flush=True)
    if a.phases:
        # second line: per-phase device time of stat_grad_shared (CUDA events recorded inside the library)
        try:
            import ctypes as C
            lib = tr.engine.lib
            lib.stat_grad_profile_enable(1)
            tr.f_grad_shared(*batch)
            n = lib.stat_grad_profile_phases()
            msv, cnt = (C.c_float * n)(), (C.c_int * n)()
            lib.stat_grad_profile_collect(msv, cnt, n)
            lib.stat_grad_profile_enable(0)
            phases = {lib.stat_grad_profile_phase_name(i).decode(): {'ms': round(float(msv[i]), 4), 'n': int(cnt[i])}
                      for i in range(n)}
        except Exception as e:
            phases = {'error': str(e)[:200]}
        print(json.dumps({'backward_phases': phases}), flush=True)


if __name__ == '__main__':
    main()
