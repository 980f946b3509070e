"""Timeline of the fused attention kernel's roles (CTA 0) at the bench shape: build the instrumented variant with
    tools/build_variant.sh fa_ts -DHMVIT_TS
and run on the GPU box with  HMVIT_LIB=build/variants/lib_fa_ts.so python tools/fa_timeline.py [kind]
Prints, per role, the recorded events in cycles relative to the CTA's first event."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import hmvit_loader
from oracle import hmvit_oracle as O
import bench

kind = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
lib, ops = pkg._lib, pkg.ops
cfg = O.default_config()
net = pkg.HeteroFusion(cfg).eval(); net.load_state_dict(O.synth_state_dict(cfg, 0)); net = net.to(dev)
B, L, C, H, W = 8, bench.L, bench.C, bench.H, bench.W
N = H * W
x, T, mode, rl, mask = bench.make_inputs(1236, B)
x, T, mode = x.to(dev), T.to(dev), mode.to(dev)
rl, cav = rl.to(torch.int32).to(dev), mask.to(torch.int32).to(dev)
blk = net.hetero_fusion_block
w = blk.packed()["window" if kind == 0 else "grid"]
rows = B * L * N
qkv = torch.empty(5, rows, C, dtype=torch.bfloat16, device=dev)
out = torch.zeros(rows, C, dtype=torch.bfloat16, device=dev)
ops.rowgemm(lib.GEMM_QKV, n_out=1280, a=x, w0=w["wqkv0"], w1=w["wqkv1"], bias=w["bqkv"], out=qkv, B=B, L=L, N=N, mode=mode, record_len=rl)
cell = float(blk.discrete_ratio) * float(blk.downsample_rate)
for _ in range(5):
    ops.group_attn(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell, q=qkv[0], k=qkv[1:3],
                   v=qkv[3:5], bk=w["bk"], bv=w["bv"], bias_table=w["bias_table"], out=out)
torch.cuda.synchronize()
so = lib.load()
buf = np.zeros((2, 4, 1024), dtype=np.uint64)
rc = so.hmvit_debug_fa_ts(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0
names = ["softmax", "gather", "mma", "qload"]
for cta in range(1):
    t0 = min(int(buf[cta, r, 0]) >> 8 for r in range(4) if buf[cta, r, 0])
    for r in range(4):
        ev = [(int(v) & 0xff, (int(v) >> 8) - t0) for v in buf[cta, r] if v]
        lo = int(os.environ.get("TS_FROM", "120"))
        print(names[r], len(ev), f"events; (code:delta-to-previous) from event {lo}:")
        print(" ".join(f"{c}:{t - ev[i - 1][1]}" for i, (c, t) in enumerate(ev) if lo <= i < lo + 170))
# merged absolute timeline of CTA 0 (TS_ABS=1): time, role, code
if os.environ.get("TS_ABS"):
    allev = []
    for r in range(4):
        allev += [((int(v) >> 8) - t0, names[r], int(v) & 0xff) for v in buf[0, r] if v]
    allev.sort()
    lo, hi = int(os.environ.get("TS_T0", "300000")), int(os.environ.get("TS_T1", "380000"))
    print("merged timeline (cycles since first event):")
    for t, n, c in allev:
        if lo <= t < hi:
            print(f"{t:8d} {n:8s} {c}")
