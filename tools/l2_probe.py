"""How much of a cyclically re-read working set does the B200 L2 keep?  x.sum() over S MB,
alone and interleaved with a streamed second buffer of W MB (the per-step weights)."""
import torch
dev = 'cuda'
def t_us(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
p = torch.cuda.get_device_properties(0)
print('L2', p.L2_cache_size / 2**20, 'MB', 'persisting max', getattr(p, 'persisting_l2_cache_max_size', None))
w = torch.empty(42 * 2**20 // 4, device=dev).normal_()
for S in (16, 32, 48, 64, 80, 96, 112, 128, 160, 256):
    x = torch.empty(S * 2**20 // 4, device=dev).normal_()
    t1 = t_us(lambda: x.sum())
    def both():
        x.sum()
        w.sum()
    tw = t_us(lambda: w.sum())
    t2 = t_us(both)
    print('S=%3d MB: alone %.1f us = %.2f TB/s | with 42 MB stream: %.1f us (stream alone %.1f us) -> x part %.2f TB/s'
          % (S, t1, S * 2**20 / t1 / 1e6, t2, tw, S * 2**20 / max(t2 - 42 * 2**20 / 6.0e6, 1e-3) / 1e6))
    del x
