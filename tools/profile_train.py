"""torch-profiler kernel table of one config-4 training step (which kernels besides hmvit::* take time):
python tools/profile_train.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import hmvit_loader
from oracle import hmvit_oracle as O
import bench

dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
cfg = O.default_config()
cfg["hetero_fusion_block"]["drop_out"] = 0.0
net = pkg.HeteroFusion(cfg).train()
net.load_state_dict(O.synth_state_dict(cfg, 0), strict=True)
net = net.to(dev)
x, T, mode, rl, mask = bench.make_inputs(1238, 8)
xd = x.to(dev).requires_grad_(True)
inp = [t.to(dev) for t in (T, mode, rl, mask)]
g = torch.randn(8, bench.C, bench.H, bench.W, device=dev)
bucket = pkg.FlatGradAllReduce(net)


def step():
    bucket.zero_grad()
    xd.grad = None
    y = net(xd, *inp)
    (y * g).sum().backward()
    bucket.allreduce(None)


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
tab = prof.key_averages()
rows = sorted(tab, key=lambda e: -e.device_time_total)
tot = sum(e.self_device_time_total for e in tab)
ours = sum(e.self_device_time_total for e in tab if "hmvit" in e.key)
print(f"device time total {tot / 1e3:.2f} ms, hmvit kernels {ours / 1e3:.2f} ms, other {((tot - ours) / 1e3):.2f} ms")
n = 0
for e in sorted(tab, key=lambda e: -e.self_device_time_total):
    if e.self_device_time_total <= 0:
        continue
    print(f"{e.self_device_time_total / 1e3:8.3f} ms  x{e.count:<4d} {e.key[:110]}")
    n += 1
    if n >= 32:
        break
