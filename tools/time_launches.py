"""Per-launch CUDA-event times of one forward (config 2), stage by stage: shows what the last (ego-only) stage costs.
Usage on the GPU box: python tools/time_launches.py"""
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
ops, lib = pkg.ops, pkg._lib
log = []
orig = {n: getattr(ops, n) for n in ("rowgemm", "group_attn", "out_ffn_chain", "ffn_head")}


def wrap(name):
    fn = orig[name]

    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record()
        log.append((name + ("/ego" if k.get("ego_only") else ""), e0, e1))
        return r
    return w


with torch.no_grad():
    for _ in range(2):
        net(*inp)
    for n in orig:
        setattr(ops, n, wrap(n))
    for rep in range(4):
        if rep == 1:
            log.clear()
        bench.kernel_breakdown(pkg, net, inp, iters=0) if False else None
        k = bench.kernel_breakdown(pkg, net, inp, iters=1)
torch.cuda.synchronize()
agg = {}
for name, e0, e1 in log:
    agg.setdefault(name, []).append(e0.elapsed_time(e1))
print(json.dumps({n: [round(min(v), 4), round(sum(v) / len(v), 4), len(v)] for n, v in agg.items()}))
