"""Repeatability of the fused attention at the bench shape: N runs per partition kind against the single-kernel form
(the single kernel shares no tile / ring / barrier code with it); prints the worst rel-L2 per kind."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hmvit_loader
from oracle import hmvit_oracle as O
import bench

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
ws = torch.empty(max(ops.attn_workspace_bytes(B, L, H, W), 256), dtype=torch.uint8, device=dev)
res = {}
for kind, kname in ((0, "window"), (1, "grid")):
    w = pk[kname]
    ops.rowgemm(lib.GEMM_QKV, n_out=1280, a=x, w0=w["wqkv0"], w1=w["wqkv1"], bias=w["bqkv"], out=qkv, B=B, L=L, N=N, mode=mode, record_len=rl)
    def run(impl, out, **kw):
        ops.group_attn(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell, q=qkv[0], k=qkv[1:3],
                       v=qkv[3:5], bk=w["bk"], bv=w["bv"], bias_table=w["bias_table"], out=out, impl=impl, **kw)
    ref = torch.zeros(rows, C, dtype=torch.bfloat16, device=dev)
    run("single", ref)
    reff = ref.float(); nref = float(reff.norm())
    worst, first = 0.0, None
    for i in range(runs):
        out = torch.zeros(rows, C, dtype=torch.bfloat16, device=dev)
        run("fused", out, workspace=ws, records_valid=(i > 0))
        torch.cuda.synchronize()
        e = float((out.float() - reff).norm()) / nref
        first = e if first is None else first
        worst = max(worst, e)
    res[kname] = {"first": round(first, 6), "worst": round(worst, 6)}
print(res)
