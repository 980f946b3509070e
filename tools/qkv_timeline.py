"""Timeline of CTA 0 of the QKV kernel (build with -DHMVIT_TS, select with HMVIT_LIB): clock64() per role.
Prints, per 128-column chunk, the epilogue warp's phases and the MMA issuer's waits, and per tile the producer's."""
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
buf = (C.c_ulonglong * (3 * 512))()
lib.hmvit_debug_qkv_ts(buf)
ts = np.array(buf[:], dtype=np.int64).reshape(3, 512)
t0 = ts[2, 0]
print("producer (tile: begin, loaded, a_empty ok, a_full arrive)")
for ti in range(8):
    print(" tile", ti, " ".join(str(int(v - t0)) for v in ts[2, ti * 4:ti * 4 + 4]))
print("chunk: epilogue [wait, acc_full, tmem ld done, staged, stored] | mma [wait acc_empty, ok, committed]")
for ci in range(50):
    e = ts[0, ci * 5:ci * 5 + 5] - t0
    m = ts[1, ci * 3:ci * 3 + 3] - t0
    print(f" {ci:3d}  epi {int(e[0]):7d} +{int(e[1]-e[0]):5d} +{int(e[2]-e[1]):5d} +{int(e[3]-e[2]):5d} +{int(e[4]-e[3]):5d} | mma {int(m[0]):7d} +{int(m[1]-m[0]):5d} +{int(m[2]-m[1]):5d}")
