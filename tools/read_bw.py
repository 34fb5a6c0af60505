import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
x = torch.randn(64, 255, 76, 76, device="cuda")
y = torch.empty_like(x)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize(); ts.append(ev[0].elapsed_time(ev[1]) * 1e3)
    return min(ts)
nb = x.numel() * 4
us = t(lambda: x.sum()); print("sum (read only): %.1f us  %.0f GB/s" % (us, nb / us / 1e3))
us = t(lambda: x.max()); print("max (read only): %.1f us  %.0f GB/s" % (us, nb / us / 1e3))
us = t(lambda: y.copy_(x)); print("copy (read+write): %.1f us  %.0f GB/s" % (us, 2 * nb / us / 1e3))
us = t(lambda: x.amax(dim=1)); print("amax over channel planes (strided like ours): %.1f us  %.0f GB/s" % (us, nb / us / 1e3))
