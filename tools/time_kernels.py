"""Time individual kernels at the bench shape (config 2) through the C-ABI; used for tuning sweeps.
HMVIT_LIB=<path to a variant .so> selects the build."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hmvit_loader
from oracle import hmvit_oracle as O
import bench

dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
cfg = O.default_config()
net = pkg.HeteroFusion(cfg).eval(); net.load_state_dict(O.synth_state_dict(cfg, 0)); net = net.to(dev)
x, T, mode, rl, mask = bench.make_inputs(1236, 8)
inp = [x.to(dev), T.to(dev), mode.to(dev), rl.to(torch.int32).to(dev), mask.to(torch.int32).to(dev)]
with torch.no_grad():
    for _ in range(2):
        net(*inp)
    k = bench.kernel_breakdown(pkg, net, inp, iters=5)
print(json.dumps({"lib": os.environ.get("HMVIT_LIB", "default"), **{n: round(v["ms_per_launch"], 4) for n, v in k.items()}}))
