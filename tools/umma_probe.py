"""Bring-up: check UMMA operand layouts (K-major / MN-major B, N = 16) against torch on the GPU."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hmvit_loader
pkg = hmvit_loader.load()
lib = pkg._lib.load()
lib.hmvit_debug_umma.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
dev = torch.device("cuda:0")
torch.manual_seed(0)
A = torch.randn(128, 64, device=dev).to(torch.bfloat16)

def run(N, mn, lbo=16, sbo=1024, kstep=2048):
    if mn & 1:
        Bm = torch.randn(64, N, device=dev).to(torch.bfloat16)      # [k][n]
        ref = A.float() @ Bm.float()
    else:
        Bm = torch.randn(N, 64, device=dev).to(torch.bfloat16)      # [n][k]
        ref = A.float() @ Bm.float().t()
    D = torch.zeros(128, N, device=dev)
    rc = lib.hmvit_debug_umma(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), N, mn, lbo, sbo, kstep, None)
    torch.cuda.synchronize()
    err = float((D - ref).abs().max() / ref.abs().max())
    if err > 1e-3:
        colerr = (D - ref).abs().amax(0) / ref.abs().max()
        rowerr = (D - ref).abs().amax(1) / ref.abs().max()
        print("   bad cols:", [i for i, v in enumerate(colerr.tolist()) if v > 1e-3][:40], "bad rows:", [i for i, v in enumerate(rowerr.tolist()) if v > 1e-3][:40])
    print(f"N={N} mn_major={mn} lbo={lbo} sbo={sbo} kstep={kstep}: rc={rc} rel max err {err:.3e}")

run(64, 0)
run(16, 2)
run(32, 2)
run(48, 2)
run(80, 1, 8192)
run(80, 3, 8192)
run(128, 1, 8192)
run(16, 3)
run(32, 3)
