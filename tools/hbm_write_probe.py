import torch
x = torch.empty(805306368 // 4, dtype=torch.float32, device="cuda")
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.zero_()); print("write-only  805 MB: %.1f us  %.2f TB/s" % (ms*1e3, 0.805/ms))
ms = t(lambda: x.fill_(1.5)); print("fill        805 MB: %.1f us  %.2f TB/s" % (ms*1e3, 0.805/ms))
ms = t(lambda: y.copy_(x)); print("copy  2x805 MB: %.1f us  %.2f TB/s" % (ms*1e3, 1.61/ms))
ms = t(lambda: x.sum()); print("read-only   805 MB: %.1f us  %.2f TB/s" % (ms*1e3, 0.805/ms))
