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
buf = (C.c_ulonglong * (8 * 3 * 64))()
lib.hmvit_debug_tc_ts(buf)
ts = np.array(buf[:], dtype=np.int64).reshape(8, 3, 64)
for cta in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    t0 = ts[cta, 2, 0]
    sm, ga = ts[cta, 0], ts[cta, 1]
    print(f"cta {cta}: staging done @{sm[0]-t0}")
    u = 0
    while 1 + u * 4 + 3 < 64 and sm[1 + u * 4 + 1] > 0:
        e = sm[1 + u * 4: 1 + u * 4 + 4] - t0
        nxt = sm[1 + (u + 1) * 4] - t0
        print(f"   softmax unit {u}: wait_S@{e[0]} got_S@{e[1]} (+{e[1]-e[0]}) p_ready@{e[2]} (+{e[2]-e[1]}) p_empty@{e[3]} (+{e[3]-e[2]}) done@{nxt} (+{nxt-e[3]})")
        u += 1
    print(f"   softmax final wait done @{sm[1 + u * 4]-t0}, stores done @{sm[2 + u * 4]-t0}")
    j = 0
    while 1 + j * 6 + 5 < 64 and ga[1 + j * 6] > 0:
        e = ga[1 + j * 6: 1 + j * 6 + 6] - t0
        print(f"   gather src {j}: start@{e[0]} Kloads_issued@{e[1]} (+{e[1]-e[0]}) k_empty@{e[2]} (+{e[2]-e[1]}) k_full@{e[3]} (+{e[3]-e[2]}) Vloads@{e[4]} (+{e[4]-e[3]}) v_empty@{e[5]} (+{e[5]-e[4]})")
        j += 1
