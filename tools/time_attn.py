"""Time the attention forms at the bench shape (config 2): single fused kernel vs the split form
(warp + compaction pass, dense attention pass), per partition; also reports their output difference.
HMVIT_LIB=<variant .so> selects the build."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hmvit_loader
from oracle import hmvit_oracle as O
import bench

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
pk = blk.packed()
rows = B * L * N
qkv = torch.empty(5, rows, C, dtype=torch.bfloat16, device=dev)
cell = float(blk.discrete_ratio) * float(blk.downsample_rate)
common = dict(B=B, L=L, N=N, mode=mode, record_len=rl)
so = lib.load()
so.hmvit_debug_split_phase.argtypes = [ctypes.c_int]


def t_ms(fn, iters=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {"lib": os.environ.get("HMVIT_LIB", "default")}
with torch.no_grad():
    for _ in range(30):                                   # clocks up before anything is timed
        net(x, T, mode, rl, cav)
    torch.cuda.synchronize()
    for kind, kname in ((0, "window"), (1, "grid")):
        w = pk[kname]
        ops.rowgemm(lib.GEMM_QKV, n_out=1280, a=x, w0=w["wqkv0"], w1=w["wqkv1"], bias=w["bqkv"], out=qkv, **common)
        for dead in (False, True):
            if dead and kind == 0:
                continue
            outs = {}

            def run(split, tag):
                out = outs.setdefault(tag, torch.zeros(rows, C, dtype=torch.bfloat16, device=dev))
                ops.group_attn(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell,
                               q=qkv[0], k=qkv[1:3], v=qkv[3:5], bk=w["bk"], bv=w["bv"], bias_table=w["bias_table"],
                               out=out, ego_only=dead, split=split)
            name = kname + ("_ego" if dead else "")
            res[name + "_single"] = round(t_ms(lambda: run(False, "single")), 4)
            res[name + "_split"] = round(t_ms(lambda: run(True, "split")), 4)
            so.hmvit_debug_split_phase(1)
            res[name + "_compact"] = round(t_ms(lambda: run(True, "tmp")), 4)
            so.hmvit_debug_split_phase(2)
            res[name + "_dense"] = round(t_ms(lambda: run(True, "tmp")), 4)
            so.hmvit_debug_split_phase(0)
            d = (outs["single"].float() - outs["split"].float())
            res[name + "_maxdiff"] = float(d.abs().max())
            res[name + "_rel_l2"] = float(d.norm() / outs["single"].float().norm())
print(json.dumps(res))
