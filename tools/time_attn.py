"""Time the attention implementations at the bench shape (config 2), per partition kind and for the ego-only stage:
fused persistent tcgen05 kernel (with valid key records), the key-record pass alone, the split form and the single
kernel; also reports the output difference fused vs single.  HMVIT_LIB=<variant .so> selects a tuning build
(tools/build_variant.sh); TIME_ATTN_IMPLS=fused[,split,single] restricts the forms."""
import json, os, sys
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
impls = os.environ.get("TIME_ATTN_IMPLS", "fused,split,single").split(",")


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
    for _ in range(20):                                   # clocks up before anything is timed
        net(x, T, mode, rl, cav)
    torch.cuda.synchronize()
    ws = torch.empty(max(ops.attn_workspace_bytes(B, L, H, W), 256), dtype=torch.uint8, device=dev)
    for kind, kname in ((0, "window"), (1, "grid")):
        w = pk[kname]
        ops.rowgemm(lib.GEMM_QKV, n_out=1280, a=x, w0=w["wqkv0"], w1=w["wqkv1"], bias=w["bqkv"], out=qkv, **common)
        for dead in (False, True):
            if dead and kind == 0:
                continue
            outs = {}

            def run(impl, tag, **kw):
                out = outs.setdefault(tag, torch.zeros(rows, C, dtype=torch.bfloat16, device=dev))
                ops.group_attn(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell,
                               q=qkv[0], k=qkv[1:3], v=qkv[3:5], bk=w["bk"], bv=w["bv"], bias_table=w["bias_table"],
                               out=out, ego_only=dead, impl=impl, **kw)
            name = kname + ("_ego" if dead else "")
            if "fused" in impls:
                run("fused", "fused", workspace=ws, records_valid=False)
                res[name + "_fused"] = round(t_ms(lambda: run("fused", "fused", workspace=ws, records_valid=True)), 4)
                res[name + "_fused+records"] = round(t_ms(lambda: run("fused", "fused", workspace=ws, records_valid=False)), 4)
            for impl in ("split", "single"):
                if impl in impls:
                    res[name + "_" + impl] = round(t_ms(lambda: run(impl, impl)), 4)
            if "fused" in impls and "single" in impls:
                d = (outs["single"].float() - outs["fused"].float())
                res[name + "_fused_vs_single_rel_l2"] = float(d.norm() / outs["single"].float().norm())
print(json.dumps(res))
