"""Timeline of the first CTAs of the tcgen05 dense attention kernel (build with -DHMVIT_TS -DHMVIT_DENSE_TC_DEFAULT=1)."""
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
net.skip_dead_queries = False
x, T, mode, rl, mask = bench.make_inputs(1236, 8)
inp = [x.to(dev), T.to(dev), mode.to(dev), rl.to(torch.int32).to(dev), mask.to(torch.int32).to(dev)]
with torch.no_grad():
    for _ in range(2):
        net(*inp)
torch.cuda.synchronize()
lib = pkg._lib.load()
buf = (C.c_ulonglong * (8 * 4 * 32))()
lib.hmvit_debug_dtc_ts(buf)
ts = np.array(buf[:], dtype=np.int64).reshape(8, 4, 32)
for cta in range(4):
    t0 = ts[cta, 0, 0]
    def rel(v): return int(v - t0) if v else -1
    print(f"cta {cta}: start 0, staged {rel(ts[cta,0,1])}, end {rel(ts[cta,0,28])}, dealloc {rel(ts[cta,2,29])}")
    for pr in range(2):
        ev = " ".join(f"[wait {rel(ts[cta,pr,2+v*3])} s_full {rel(ts[cta,pr,3+v*3])} p_out {rel(ts[cta,pr,4+v*3])}]" for v in range(5) if ts[cta, pr, 3 + v * 3])
        print(f"   pair {pr}: {ev} loop_end {rel(ts[cta,pr,26])} stored {rel(ts[cta,pr,27])}")
