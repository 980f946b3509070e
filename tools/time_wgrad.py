"""Per-launch time of the typed weight-gradient kernel (hmvit_bwd_wgrad) and of hmvit_bwd_dgrad_cat at the bench shape
(8 scenes x 5 agents x 48 x 176), one line per operand combination with its algorithmic bytes and the HBM rate they imply.
Usage on the GPU box: python tools/time_wgrad.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import hmvit_loader  # noqa: E402

dev = torch.device("cuda:0")
pkg = hmvit_loader.load()
ops = pkg.ops
B, L, N = 8, 5, 48 * 176
R = B * L * N
g = torch.Generator().manual_seed(0)
mode = (torch.rand(B, L, generator=g) < 0.5).to(torch.int32).to(dev)
rl = torch.full((B,), L, dtype=torch.int32, device=dev)
a_cm = torch.randn(B * L, 256, N, device=dev)
b_cm = torch.randn(B * L, 256, N, device=dev)
rows = torch.randn(5, R, 256, device=dev).to(torch.bfloat16)
st = torch.zeros(R, 2, device=dev)
ops.bwd_row_stats(b_cm, st, B=B, L=L, N=N, record_len=rl)
dw = torch.zeros(2, 1280, 256, device=dev)
w = [torch.randn(1280, 256, device=dev).to(torch.bfloat16) for _ in range(2)]
out = torch.empty(B * L, 256, N, device=dev)
kw = dict(B=B, L=L, N=N, mode=mode, record_len=rl)
MB = 1e6
cases = {
    "wgrad rows x cm(+LN)": (lambda: ops.bwd_wgrad(rows[0], b_cm, dw, b_stats=st, row0=256, **kw), (R * 256 * 2 + R * 256 * 4 + R * 8) / MB),
    "wgrad cm x cm": (lambda: ops.bwd_wgrad(a_cm, b_cm, dw, **kw), 2 * R * 256 * 4 / MB),
    "wgrad cm x cm(+LN)": (lambda: ops.bwd_wgrad(a_cm, b_cm, dw, b_stats=st, **kw), (2 * R * 256 * 4 + R * 8) / MB),
    "wgrad cm x rows": (lambda: ops.bwd_wgrad(a_cm, rows[1], dw, **kw), (R * 256 * 2 + R * 256 * 4) / MB),
    "wgrad rows x cm, ego only": (lambda: ops.bwd_wgrad(rows[0], b_cm, dw, ego_only=True, **kw), (R * 256 * 6 / L) / MB),
    "dgrad_cat (K = 1280)": (lambda: ops.bwd_dgrad_cat(rows, w[0], w[1], out, **kw), (5 * R * 256 * 2 + R * 256 * 4) / MB),
}
for name, (fn, mb) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"case": name, "ms_per_launch": round(ms, 4), "algorithmic_MB": round(mb, 1), "GB_per_s": round(mb / ms, 1)}))
