"""Diagnose H2D / compute overlap: times the pinned->device copy alone, the forward alone, and both
issued concurrently on different streams."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hmvit_loader
from oracle import hmvit_oracle as O

dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
cfg = O.default_config()
net = pkg.HeteroFusion(cfg).eval(); net.load_state_dict(O.synth_state_dict(cfg, 0)); net = net.to(dev)
x, T, mode, rl, mask = O.synth_inputs(8, 5, 256, 48, 176, [5] * 8, 1)
hx = x.pin_memory()
print("pinned:", hx.is_pinned())
dx = [torch.empty_like(x, device=dev) for _ in range(2)]
inp = [t.to(dev) for t in (x, T, mode, rl, mask)]
s2 = torch.cuda.Stream()

def ev():
    return torch.cuda.Event(enable_timing=True)

with torch.no_grad():
    for _ in range(3):
        net(*inp)
    torch.cuda.synchronize()
    # copy alone
    a, b = ev(), ev()
    with torch.cuda.stream(s2):
        a.record(s2)
        for k in range(5):
            dx[k & 1].copy_(hx, non_blocking=True)
        b.record(s2)
    torch.cuda.synchronize()
    print("copy alone  ms/step", a.elapsed_time(b) / 5, "GB/s", hx.numel() * 4 / (a.elapsed_time(b) / 5 * 1e-3) / 1e9)
    # compute alone
    a, b = ev(), ev()
    a.record()
    for k in range(5):
        net(*inp)
    b.record()
    torch.cuda.synchronize()
    print("compute alone ms/step", a.elapsed_time(b) / 5)
    # concurrent
    t0 = time.perf_counter()
    a, b, c, d = ev(), ev(), ev(), ev()
    a.record()
    with torch.cuda.stream(s2):
        c.record(s2)
        for k in range(5):
            dx[k & 1].copy_(hx, non_blocking=True)
        d.record(s2)
    t1 = time.perf_counter()
    for k in range(5):
        net(*inp)
    b.record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("concurrent: compute ms/step", a.elapsed_time(b) / 5, "copy ms/step", c.elapsed_time(d) / 5,
          "cpu enqueue copy %.2f ms, enqueue compute %.2f ms, wall %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t0) * 1e3))
