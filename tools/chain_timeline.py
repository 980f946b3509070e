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
with torch.no_grad():
    for _ in range(2):
        net(*inp)
torch.cuda.synchronize()
lib = pkg._lib.load()
buf = (C.c_ulonglong * (2 * 16 * 16))()
lib.hmvit_debug_chain_ts(buf)
ts = np.array(buf[:], dtype=np.int64).reshape(2, 16, 16)
t0 = ts[1, 0, 0]
names_t = ["E1wait", "d1_full", "E1done", "P2fed", "d2_full", "P3fed", "d1_final", "E2done", "stats"]
names_m = ["loop", "d1_free", "o_full", "P1iss", "P2iss", "P3iss"]
for tile in range(6):
    print("tile", tile, "| transform:", " ".join(f"{n}={int(ts[0,tile,i]-t0)}" for i, n in enumerate(names_t)))
    print("        | mma      :", " ".join(f"{n}={int(ts[1,tile,i]-t0)}" for i, n in enumerate(names_m)))
