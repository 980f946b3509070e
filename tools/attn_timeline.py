import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, numpy as np
import hmvit_loader
from oracle import hmvit_oracle as O
import bench
dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
cfg = O.default_config()
net = pkg.HeteroFusion(cfg).eval(); net.load_state_dict(O.synth_state_dict(cfg, 0)); net = net.to(dev)
x, T, mode, rl, mask = bench.make_inputs(1236, 8)
inp = [x.to(dev), T.to(dev), mode.to(dev), rl.to(torch.int32).to(dev), mask.to(torch.int32).to(dev)]
net.skip_dead_queries = False
with torch.no_grad():
    net(*inp)
torch.cuda.synchronize()
lib = pkg._lib.load()
buf = (C.c_ulonglong * (8 * 8 * 4))()
lib.hmvit_debug_attn_ts(buf)
ts = np.array(buf[:], dtype=np.int64).reshape(8, 8, 4)
for cta in range(8):
    base = ts[cta][ts[cta] > 0].min() if (ts[cta] > 0).any() else 0
    out = [f"taps@{int(ts[cta,0,0]-base)}"]
    for j in range(5):
        e = ts[cta, j]
        if e[1] == 0: continue
        out.append(f"j{j}: start@{int(e[1]-base)} gather+{int(e[2]-e[1])} compute+{int(e[3]-e[2])}")
    print("cta", cta, " | ".join(out))
