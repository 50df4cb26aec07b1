"""Host-to-device bandwidth for one bench batch (259 MB of pinned fp32 features): one tensor, the four
feature tensors on one stream, the big one split over two streams."""
import torch
n = 259004928 // 4
big = torch.empty(n, dtype=torch.float32).pin_memory()
parts = [torch.empty(s, dtype=torch.float32).pin_memory() for s in (64 * 26 * 2048, 64 * 26, 64 * 26 * 8 * 4096, 64 * 26 * 4096)]
dbig = torch.empty(n, dtype=torch.float32, device='cuda')
dparts = [torch.empty_like(p, device='cuda') for p in parts]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def one():
    dbig.copy_(big, non_blocking=True)


def four():
    for d, p in zip(dparts, parts):
        d.copy_(p, non_blocking=True)


def two_streams():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    h = parts[2].numel() // 2
    with torch.cuda.stream(s1):
        dparts[0].copy_(parts[0], non_blocking=True)
        dparts[2][:h].copy_(parts[2][:h], non_blocking=True)
    with torch.cuda.stream(s2):
        dparts[3].copy_(parts[3], non_blocking=True)
        dparts[2][h:].copy_(parts[2][h:], non_blocking=True)
    cur.wait_stream(s1)
    cur.wait_stream(s2)


for name, fn in (('one 259 MB copy', one), ('four tensors, one stream', four), ('two streams', two_streams)):
    ms = timed(fn)
    print('%-26s %.3f ms  %.1f GB/s' % (name, ms, 259004928 / ms / 1e6))
