"""Where does the fused attention differ from the single-kernel form?  Per head / per agent / per group error map at the bench shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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
B, L, C, H, W = 2, bench.L, bench.C, bench.H, bench.W
N = H * W
x, T, mode, rl, mask = bench.make_inputs(1236, B)
x, T, mode = x.to(dev), T.to(dev), mode.to(dev)
rl, cav = rl.to(torch.int32).to(dev), mask.to(torch.int32).to(dev)
blk = net.hetero_fusion_block
w = blk.packed()["window" if kind == 0 else "grid"]
rows = B * L * N
qkv = torch.empty(5, rows, C, dtype=torch.bfloat16, device=dev)
ops.rowgemm(lib.GEMM_QKV, n_out=1280, a=x, w0=w["wqkv0"], w1=w["wqkv1"], bias=w["bqkv"], out=qkv, B=B, L=L, N=N, mode=mode, record_len=rl)
cell = float(blk.discrete_ratio) * float(blk.downsample_rate)
outs = {}
ws = torch.empty(max(ops.attn_workspace_bytes(B, L, H, W), 256), dtype=torch.uint8, device=dev)
for impl in ("fused", "single"):
    out = torch.zeros(rows, C, dtype=torch.bfloat16, device=dev)
    lse = torch.zeros(rows, 8, dtype=torch.float32, device=dev)
    kw = dict(workspace=ws, records_valid=False) if impl == "fused" else {}
    kw["lse"] = lse
    ops.group_attn(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell, q=qkv[0], k=qkv[1:3],
                   v=qkv[3:5], bk=w["bk"], bv=w["bv"], bias_table=w["bias_table"], out=out, impl=impl, **kw)
    torch.cuda.synchronize()
    outs[impl] = out.float().view(B * L, H, W, 8, 32)
    outs[impl + "_lse"] = lse.view(B * L, H, W, 8)
d = outs["fused"] - outs["single"]
ref = outs["single"]
print("total rel-L2", float(d.norm() / ref.norm()))
print("per head", [round(float(d[..., h, :].norm() / ref[..., h, :].norm()), 5) for h in range(8)])
print("per agent", [round(float(d[a].norm() / (ref[a].norm() + 1e-9)), 5) for a in range(B * L)])
e = d.pow(2).sum(dim=(3, 4)).sqrt() / (ref.pow(2).sum(dim=(3, 4)).sqrt() + 1e-9)    # [BL, H, W]
print("rows with rel err > 0.02:", int((e > 0.02).sum()), "of", e.numel(), " max", float(e.max()))
bad = (e > 0.02).nonzero()[:20]
print(bad.tolist())
wb = (e > 0.02).view(B * L, H // 8, 8, W // 8, 8).any(dim=4).any(dim=2)      # [BL, GY, GX]
idx = wb.nonzero().tolist()
G = (H // 8) * (W // 8)
nvis = torch.frombuffer(ws.cpu().numpy().tobytes(), dtype=torch.int32)
recs_bytes = B * L * G * L * 64 * 16
nv = torch.frombuffer(ws[recs_bytes:recs_bytes + B * L * G * 4].cpu().numpy().tobytes(), dtype=torch.int32).view(B * L, G)
valid = [a for a in range(B * L) if (a % L) < int(rl[a // L])]
print("bad windows:", len(idx))
for a, gy, gx in idx[:40]:
    grp = gy * (W // 8) + gx if kind == 0 else None
    it = valid.index(a) * G + (gy * (W // 8) + gx)
    step = 74
    print(f"agent {a} win ({gy},{gx}) item {it} cta {it % step} k {it // step} nv {int(nv[a, gy * (W // 8) + gx])} prev_nv {int(nv.view(-1)[[valid[(it - step) // G] * G + (it - step) % G]][0]) if it >= step else -1}")
dh = d.view(B * L, H // 8, 8, W // 8, 8, 8, 32)
rh = ref.view(B * L, H // 8, 8, W // 8, 8, 8, 32)
for a, gy, gx in idx[:12]:
    errs = [round(float(dh[a, gy, :, gx, :, h].norm() / (rh[a, gy, :, gx, :, h].norm() + 1e-9)), 3) for h in range(8)]
    rows_bad = int((e.view(B * L, H // 8, 8, W // 8, 8)[a, gy, :, gx, :] > 0.02).sum())
    print(f"agent {a} win ({gy},{gx}) per-head rel err {errs} bad rows {rows_bad}")

dl = (outs["fused_lse"] - outs["single_lse"]).abs()
dl = torch.where(torch.isfinite(dl), dl, torch.zeros_like(dl))
print("lse abs diff: max", float(dl.max()), "mean", float(dl.mean()))
dlw = dl.view(B * L, H // 8, 8, W // 8, 8, 8)
for a, gy, gx in idx[:12]:
    print(f"agent {a} win ({gy},{gx}) lse diff per head (max over rows)", [round(float(dlw[a, gy, :, gx, :, h].max()), 3) for h in range(8)])
