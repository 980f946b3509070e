"""GPU parity checks shared by the pytest suite (-m gpu) and tools/bringup.py.

Each check runs the CUDA path through the C-ABI (via hm-vit_b200/ops.py or the nn.Module surface)
and compares with the CPU oracle / emulation on the same seeded inputs.  Returns a dict of error
metrics; raises AssertionError when outside the stated tolerance.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import hmvit_loader  # noqa: E402
from oracle import hmvit_emul as E  # noqa: E402
from oracle import hmvit_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DEV = "cuda:0"


def pkg():
    return hmvit_loader.load()


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def tf32(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def bf16(t):
    return t.to(torch.bfloat16).float()


def _mk_module(seed=0):
    cfg = O.default_config()
    P = O.synth_state_dict(cfg, seed)
    net = pkg().HeteroFusion(cfg).eval()
    net.load_state_dict(P, strict=True)
    return cfg, P, net.to(DEV)


def _scene(B, L, H, W, record_len, seed, **kw):
    return O.synth_inputs(B, L, 256, H, W, record_len, seed, **kw)


# ----------------------------------------------------------------------------------------------
def check_rowgemm(variant_name, N=200, seed=3):
    """One row-GEMM variant against a torch-CPU fp32 evaluation with the same operand rounding."""
    p = pkg()
    lib = p._lib
    B, L, C = 2, 3, 256
    g = torch.Generator().manual_seed(seed)
    record_len = torch.tensor([3, 2], dtype=torch.int32)
    mode = torch.tensor([[1, 0, 1], [0, 0, 0]], dtype=torch.int32)
    x = torch.randn(B, L, C, N, generator=g) * 1.5 + 0.3
    resid = torch.randn(B, L, C, N, generator=g)
    gam = 1 + 0.1 * torch.randn(2, C, generator=g)
    bet = 0.1 * torch.randn(2, C, generator=g)
    n_out = 1280 if variant_name.startswith("QKV") else 256
    w = torch.randn(2, n_out, C, generator=g) / 16
    bias = 0.1 * torch.randn(2, n_out, generator=g)
    d = lambda t: t.to(DEV).contiguous()
    ego_only = variant_name.endswith("_EGO")
    base = variant_name.replace("_EGO", "")
    variant = getattr(lib, "GEMM_" + base)
    ln = lambda a, t: F.layer_norm(a, (C,), gam[t], bet[t], 1e-5)
    res = {}
    if base in ("QKV", "QKV_NOLN"):
        wq = bf16(w)
        out = torch.full((5, B * L * N, C), float("nan"), dtype=torch.bfloat16, device=DEV)
        p.ops.rowgemm(variant, B=B, L=L, N=N, n_out=n_out, mode=d(mode), record_len=d(record_len), a=d(x),
                      w0=d(w[0].to(torch.bfloat16)), w1=d(w[1].to(torch.bfloat16)), bias=d(bias), out=out,
                      ln_gamma=d(gam), ln_beta=d(bet), ego_only=ego_only)
        torch.cuda.synchronize()
        out = out.float().cpu().view(5, B, L, N, C)
        worst = 0.0
        for b in range(B):
            n = int(record_len[b])
            types = [int(v) for v in mode[b, :n]]
            etypes = {types[0]} if ego_only else set(types)
            for l in range(L):
                for var in range(5):
                    te = (var - 1) % 2
                    need = l < n and ((var == 0 and (not ego_only or l == 0)) or (var > 0 and te in etypes))
                    got = out[var, b, l]
                    if not need:
                        assert torch.isnan(got).all(), f"variant {var} of (b={b}, l={l}) must not be written"
                        continue
                    t = types[l]
                    a = x[b, l].t()
                    a = bf16(ln(a, t)) if base == "QKV" else bf16(a)
                    exp = bf16(a @ wq[t, var * C:(var + 1) * C].t() + bias[t, var * C:(var + 1) * C])
                    assert torch.isfinite(got).all()
                    worst = max(worst, rel_l2(got, exp))
        res["rel_l2"] = worst
        assert worst < 4e-3, res            # output is bf16: one-ulp flips at rounding boundaries only
        return res
    # fp32 channel-major outputs
    out = torch.full((B, 1 if base == "HEAD2" else L, C, N), float("nan"), device=DEV)
    if base == "OUT":
        a_rows = torch.randn(B * L * N, C, generator=g).to(torch.bfloat16)
        p.ops.rowgemm(variant, B=B, L=L, N=N, n_out=256, mode=d(mode), record_len=d(record_len), a=d(a_rows),
                      w0=d(w[0].to(torch.bfloat16)), w1=d(w[1].to(torch.bfloat16)), bias=d(bias), out=out,
                      resid=d(resid), ego_only=ego_only)
    else:
        p.ops.rowgemm(variant, B=B, L=L, N=N, n_out=256, mode=d(mode), record_len=d(record_len), a=d(x),
                      w0=d(tf32(w[0])), w1=d(tf32(w[1])), bias=d(bias), out=out, ln_gamma=d(gam), ln_beta=d(bet),
                      resid=d(resid), ego_only=ego_only)
    torch.cuda.synchronize()
    out = out.cpu()
    worst = 0.0
    for b in range(B):
        n = int(record_len[b])
        for l in range(L):
            t = int(mode[b, l])
            slot_only0 = ego_only or base.startswith("HEAD")
            need = l < n and (not slot_only0 or l == 0)
            if base == "HEAD2":
                if l != 0:
                    continue
                got = out[b, 0]
            else:
                got = out[b, l]
            if not need:
                assert torch.isnan(got).all(), f"(b={b}, l={l}) must not be written"
                continue
            if base == "OUT":
                exp = a_rows.float().view(B, L, N, C)[b, l] @ bf16(w[t]).t() + bias[t] + resid[b, l].t()
            elif base == "FFN1":
                exp = tf32(F.gelu(tf32(ln(x[b, l].t(), t)) @ tf32(w[t]).t() + bias[t]))
            elif base == "FFN2":
                exp = tf32(x[b, l].t().contiguous()) @ tf32(w[t]).t() + bias[t] + resid[b, l].t()
            elif base == "HEAD1":
                exp = tf32(F.gelu(tf32(x[b, l].t().contiguous()) @ tf32(w[t]).t() + bias[t]))
            else:  # HEAD2
                exp = tf32(x[b, l].t().contiguous()) @ tf32(w[t]).t() + bias[t]
            assert torch.isfinite(got).all()
            worst = max(worst, rel_l2(got.t(), exp))
    res["rel_l2"] = worst
    assert worst < 2e-4, res
    return res


def check_chain(N=200, ego_only=False, seed=9):
    """Fused OUT+FFN chain kernel against a torch-CPU evaluation with the same operand rounding and
    against the three-kernel row-GEMM sequence."""
    p = pkg()
    lib = p._lib
    B, L, C = 2, 3, 256
    g = torch.Generator().manual_seed(seed)
    record_len = torch.tensor([3, 2], dtype=torch.int32)
    mode = torch.tensor([[1, 0, 1], [0, 1, 0]], dtype=torch.int32)
    o = torch.randn(B * L * N, C, generator=g).to(torch.bfloat16)
    x = torch.randn(B, L, C, N, generator=g) * 1.3 + 0.2
    gam = 1 + 0.1 * torch.randn(2, C, generator=g)
    bet = 0.1 * torch.randn(2, C, generator=g)
    wa = torch.randn(2, C, C, generator=g) / 16
    w1 = torch.randn(2, C, C, generator=g) / 16
    w2 = torch.randn(2, C, C, generator=g) / 16
    ba, b1, b2 = (0.1 * torch.randn(2, C, generator=g) for _ in range(3))
    d = lambda t: t.to(DEV).contiguous()
    wa_d = [d(wa[t].to(torch.bfloat16)) for t in range(2)]
    w1_d = [d(tf32(w1[t])) for t in range(2)]
    w2_d = [d(tf32(w2[t])) for t in range(2)]
    common = dict(B=B, L=L, N=N, mode=d(mode), record_len=d(record_len))
    out = torch.full((B, L, C, N), float("nan"), device=DEV)
    p.ops.out_ffn_chain(o=d(o), resid=d(x), out=out, wa0=wa_d[0], wa1=wa_d[1], ba=d(ba), ln_gamma=d(gam), ln_beta=d(bet),
                        w1_0=w1_d[0], w1_1=w1_d[1], b1=d(b1), w2_0=w2_d[0], w2_1=w2_d[1], b2=d(b2), ego_only=ego_only, **common)
    torch.cuda.synchronize()
    # the same arithmetic as three row-GEMM launches
    x1 = torch.full((B, L, C, N), float("nan"), device=DEV)
    hid = torch.full((B, L, C, N), float("nan"), device=DEV)
    p.ops.rowgemm(lib.GEMM_OUT, n_out=256, a=d(o), w0=wa_d[0], w1=wa_d[1], bias=d(ba), out=x1, resid=d(x), ego_only=ego_only, **common)
    p.ops.rowgemm(lib.GEMM_FFN1, n_out=256, a=x1, w0=w1_d[0], w1=w1_d[1], bias=d(b1), out=hid, ln_gamma=d(gam), ln_beta=d(bet),
                  ego_only=ego_only, **common)
    p.ops.rowgemm(lib.GEMM_FFN2, n_out=256, a=hid, w0=w2_d[0], w1=w2_d[1], bias=d(b2), out=x1, resid=x1, ego_only=ego_only, **common)
    torch.cuda.synchronize()
    out, x1 = out.cpu(), x1.cpu()
    worst, worst_seq = 0.0, 0.0
    for b in range(B):
        for l in range(L):
            need = l < int(record_len[b]) and (not ego_only or l == 0)
            if not need:
                assert torch.isnan(out[b, l]).all(), f"(b={b}, l={l}) must not be written"
                continue
            t = int(mode[b, l])
            xp = o.float().view(B, L, N, C)[b, l] @ bf16(wa[t]).t() + ba[t] + x[b, l].t()
            h = tf32(F.gelu(tf32(F.layer_norm(xp, (C,), gam[t], bet[t], 1e-5)) @ tf32(w1[t]).t() + b1[t]))
            exp = xp + h @ tf32(w2[t]).t() + b2[t]
            assert torch.isfinite(out[b, l]).all()
            worst = max(worst, rel_l2(out[b, l].t(), exp))
            worst_seq = max(worst_seq, rel_l2(out[b, l], x1[b, l]))
    res = {"rel_l2_vs_cpu": worst, "rel_l2_vs_rowgemm_sequence": worst_seq}
    assert worst < 2e-4 and worst_seq < 2e-5, res
    return res


# ----------------------------------------------------------------------------------------------
def check_warp_mask_golden():
    """Stand-alone warp and ROI mask against the reference's own outputs (tests/golden/warp_mask.npz)."""
    p = pkg()
    g = np.load(os.path.join(GOLDEN, "warp_mask.npz"))
    B, L, C, H, W = 2, 4, 2, 48, 176
    x, T, mode, record_len, mask = O.synth_inputs(B, L, C, H, W, [4, 3], 31)
    warped = torch.from_numpy(g["warped"])
    shape = tuple(int(v) for v in g["mask_shape"])
    mask_pair = torch.from_numpy(np.unpackbits(g["mask_pair"])[: int(np.prod(shape))].reshape(shape)).float()
    st = p.SpatialTransformation({"voxel_size": [0.4, 0.4, 4], "downsample_rate": 4})
    mism, worst, mism_oracle = 0, 0.0, 0
    for i in range(L):
        y = st(x.to(DEV), T[:, :, i].contiguous().to(DEV)).cpu()
        worst = max(worst, float((y - warped[:, :, i]).abs().max()))
        assert rel_l2(y, warped[:, :, i]) < 1e-4       # reference's own fp32 chain is ~1e-4 abs off closed form
        yo = O.spatial_transformation(x, T[:, :, i], 0.4, 4)
        assert rel_l2(y, yo) < 1e-6                    # vs oracle: same closed form, fp64 coordinates
        m = p.get_roi_and_cav_mask((B, L, H, W, C), mask.to(DEV), T[:, :, i].contiguous().to(DEV), 0.4, 4).cpu()
        mism += int((m != mask_pair[..., i]).sum())
        mism_oracle += int((m != O.roi_and_cav_mask((B, L, H, W, C), mask, T[:, :, i], 0.4, 4)).sum())
    res = {"warp_max_abs_vs_reference": worst, "mask_mismatch_vs_reference": mism, "mask_mismatch_vs_oracle": mism_oracle}
    assert mism == 0 and mism_oracle == 0 and worst < 5e-4, res
    return res


def check_mask_adversarial():
    """Axis-aligned poses with half-cell offsets (rounding ties): counted, compared with the oracle."""
    p = pkg()
    B, L, H, W = 1, 8, 48, 176
    T = torch.eye(4).repeat(B, L, 1, 1)
    import math
    k = 0
    for yaw in (0.0, math.pi / 2, math.pi, -math.pi / 2):
        for tx in (0.8, 2.4):
            c, s = math.cos(yaw), math.sin(yaw)
            T[0, k, 0, 0], T[0, k, 0, 1], T[0, k, 1, 0], T[0, k, 1, 1] = c, -s, s, c
            T[0, k, 0, 3], T[0, k, 1, 3] = tx, -tx
            k += 1
    cav = torch.ones(B, L, dtype=torch.int32)
    m = p.get_roi_and_cav_mask((B, L, H, W, 1), cav.to(DEV), T.to(DEV), 0.4, 4).cpu()
    mo = O.roi_and_cav_mask((B, L, H, W, 1), cav, T, 0.4, 4)
    res = {"tie_pose_mismatch_vs_oracle": int((m != mo).sum()), "pixels": int(m.numel())}
    assert res["tie_pose_mismatch_vs_oracle"] == 0, res
    # the same pose families against the REFERENCE's own masks (tests/golden/mask_ties.npz, made by
    # tests/golden/make_golden_ties.py): whole-cell shifts bit-exact; half-cell shifts (every coordinate an exact
    # rounding tie, decided by the reference's fp32 round-off) counted -- see test_oracle_golden.py for the tie proof
    g = np.load(os.path.join(GOLDEN, "mask_ties.npz"))
    n = int(np.prod(g["shape"]))
    for name in ("whole_cell", "half_cell"):
        Tg = torch.from_numpy(g[name + "_T"])
        ref = torch.from_numpy(np.unpackbits(g[name])[:n].reshape(B, H, W, 1, L)).float()
        mg = p.get_roi_and_cav_mask((B, L, H, W, 1), cav.to(DEV), Tg.to(DEV), 0.4, 4).cpu()
        assert int((mg != O.roi_and_cav_mask((B, L, H, W, 1), cav, Tg, 0.4, 4)).sum()) == 0, name
        res[name + "_mismatch_vs_reference"] = int((mg != ref).sum())
    assert res["whole_cell_mismatch_vs_reference"] == 0 and res["half_cell_mismatch_vs_reference"] <= n // 100, res
    return res


# ----------------------------------------------------------------------------------------------
def check_attention_golden():
    """Unit-level HeteroAttention.forward against the reference's output (tests/golden/attention.npz)."""
    p = pkg()
    g = np.load(os.path.join(GOLDEN, "attention.npz"))
    C, b, l, X, Y, w = 256, 2, 3, 2, 2, 8
    gen = torch.Generator().manual_seed(21)
    P = O.synth_state_dict(O.default_config(input_dim=C), 21)
    pfx = "hetero_fusion_block.grid_attention"
    att = p.HeteroAttention(C, 32, 0.1, 5, w).eval()
    att.load_state_dict({k[len(pfx) + 1:]: v for k, v in P.items() if k.startswith(pfx + ".")}, strict=True)
    att = att.to(DEV)
    x = torch.randn(b, l, X, Y, w, w, C, generator=gen)
    mode = torch.from_numpy(g["mode"])
    mask = torch.from_numpy(g["mask"]).float()
    ref = torch.from_numpy(g["out"])
    with torch.no_grad():
        y = att(x.to(DEV), mode.to(DEV), mask=mask.to(DEV)).cpu()
    res = {"rel_l2_vs_reference": rel_l2(y, ref), "max_rel": max_rel(y, ref)}
    # bf16 operands on q/k/v/p and the output projection: stated tolerance 1e-2 rel-L2 for this
    # isolated module (no fp32 residual stream around it); fused-feature tolerance is checked elsewhere.
    assert res["rel_l2_vs_reference"] < 1e-2, res
    return res


# ----------------------------------------------------------------------------------------------
def _fusion_case(B, L, H, W, record_len, seed, pseed=0, mode=None, **kw):
    cfg, P, net = _mk_module(pseed)
    x, T, md, rl, mask = _scene(B, L, H, W, record_len, seed, mode=mode, **kw)
    with torch.no_grad():
        y = net(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    torch.cuda.synchronize()
    return cfg, P, (x, T, md, rl, mask), y, net


def check_fusion_repeatable():
    """The whole forward is BIT-REPRODUCIBLE over repeated runs at the BASELINE map size (the fused attention is a
    persistent pipeline of five warp roles over mbarriers: a protocol error shows up as rare differing rows, not as a
    crash), and its fused attention agrees with the independently written mma.sync kernel at operand-rounding level
    (same bf16 operands, different tiling: measured 1.8e-3 .. 2.1e-3, bar 4e-3)."""
    cfg, P, net = _mk_module(0)
    x, T, md, rl, mask = _scene(2, 5, 48, 176, [5, 4], 77)
    inp = [t.to(DEV) for t in (x, T, md, rl, mask)]
    with torch.no_grad():
        y0 = net(*inp).clone()
        n_diff = 0
        for _ in range(8):
            n_diff += int((net(*inp) != y0).sum())
        net.hetero_fusion_block.attn_impl = "single"
        y1 = net(*inp)
    res = {"differing_elements_over_8_runs": n_diff, "fused_vs_single_rel_l2": rel_l2(y0.cpu(), y1.cpu())}
    assert n_diff == 0 and res["fused_vs_single_rel_l2"] < 4e-3, res
    return res


def check_fusion_small():
    """Whole forward, small shape, against the fp32 oracle and the operand-rounded emulation."""
    cfg, P, inp, y, net = _fusion_case(2, 3, 16, 24, [3, 2], seed=5, tx=10, ty=5)
    ref = O.hetero_fusion(*inp, P, cfg)
    res = {"rel_l2_vs_oracle": rel_l2(y, ref), "max_rel_vs_oracle": max_rel(y, ref)}
    assert torch.isfinite(y).all()
    assert res["rel_l2_vs_oracle"] < 1e-3, res
    return res


def check_fusion_golden():
    """Whole forward against the committed outputs of the reference itself."""
    g = np.load(os.path.join(GOLDEN, "fusion_c256.npz"))
    C, B, L, H, W, seed = (int(v) for v in g["meta"][:6])
    tx, ty = float(g["meta"][6]), float(g["meta"][7])
    cfg = O.default_config(input_dim=C)
    P = O.synth_state_dict(cfg, seed)
    x, T, mode, rl, mask = O.synth_inputs(B, L, C, H, W, g["record_len"].tolist(), seed + 100, tx=tx, ty=ty)
    p = pkg()
    net = p.HeteroFusion(cfg).eval()
    net.load_state_dict(P, strict=True)
    net = net.to(DEV)
    with torch.no_grad():
        y = net(x.to(DEV), T.to(DEV), mode.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
        blk = net.hetero_fusion_block(x.to(DEV), T.to(DEV), mode.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    res = {"fused_rel_l2_vs_reference": rel_l2(y, torch.from_numpy(g["fused"]))}
    ref_blk = torch.from_numpy(g["block"])
    worst = 0.0
    for b in range(B):
        n = int(rl[b])
        worst = max(worst, rel_l2(blk[b, :n], ref_blk[b, :n]))       # valid slots only
        assert torch.equal(blk[b, n:], x[b, n:])                      # padded slots pass through
    res["block_rel_l2_vs_reference"] = worst
    assert res["fused_rel_l2_vs_reference"] < 1e-3 and worst < 1e-3, res
    return res


def check_logits_golden():
    """Detection-head logits: CUDA fused feature -> decoder restatement (CPU, eval-mode BN) against the
    reference pipeline's psm / rm (tests/golden/logits_c256.npz).  Tolerance 1e-3 rel-L2 per tensor."""
    g = np.load(os.path.join(GOLDEN, "fusion_c256.npz"))
    gl = np.load(os.path.join(GOLDEN, "logits_c256.npz"))
    C, B, L, H, W, seed = (int(v) for v in g["meta"][:6])
    cfg = O.default_config(input_dim=C)
    P = O.synth_state_dict(cfg, seed)
    PD = O.synth_decoder_state_dict(seed + 1)
    x, T, mode, rl, mask = O.synth_inputs(B, L, C, H, W, g["record_len"].tolist(), seed + 100,
                                          tx=float(g["meta"][6]), ty=float(g["meta"][7]))
    net = pkg().HeteroFusion(cfg).eval()
    net.load_state_dict(P, strict=True)
    net = net.to(DEV)
    with torch.no_grad():
        y = net(x.to(DEV), T.to(DEV), mode.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    psm, rm = O.hetero_decoder(y, mode[:, 0], PD)
    res = {"psm_rel_l2": rel_l2(psm, torch.from_numpy(gl["psm"])), "rm_rel_l2": rel_l2(rm, torch.from_numpy(gl["rm"]))}
    assert res["psm_rel_l2"] < 1e-3 and res["rm_rel_l2"] < 1e-3, res
    return res


def check_decoder_vs_oracle():
    """HeteroDecoder on the GPU (TMA-shifted implicit-GEMM convolutions, fp16 operands) against the fp32 restatement of
    hetero_decoder.py:42-74 / naive_decoder.py:63-92 on a random ego feature at the BASELINE map size, mixed ego
    modalities (both weight sets in one launch).  Stated tolerance: rel-L2 <= 1e-3 per tensor."""
    torch.manual_seed(7)
    B, H, W = 3, 48, 176
    PD = O.synth_decoder_state_dict(11)
    x = torch.randn(B, 256, H, W)
    mode = torch.tensor([[1, 0], [0, 1], [1, 1]], dtype=torch.int32)
    dec = pkg().HeteroDecoder({"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}).eval()
    dec.load_state_dict(PD, strict=True)
    dec = dec.to(DEV)
    with torch.no_grad():
        psm, rm = dec(x.to(DEV).unsqueeze(1), mode.to(DEV), use_upsample=False)
    rp, rr = O.hetero_decoder(x, mode[:, 0], PD)
    res = {"psm_rel_l2": rel_l2(psm.cpu(), rp), "rm_rel_l2": rel_l2(rm.cpu(), rr),
           "psm_max_rel": max_rel(psm.cpu(), rp), "rm_max_rel": max_rel(rm.cpu(), rr)}
    assert res["psm_rel_l2"] < 1e-3 and res["rm_rel_l2"] < 1e-3, res
    # a map whose width is not a multiple of the 16-pixel tile: clipped tiles, zero padding through the TMA bounds
    x2 = torch.randn(1, 256, 16, 24)
    with torch.no_grad():
        psm2, rm2 = dec(x2.to(DEV), mode[:1].to(DEV), use_upsample=False)
    rp2, rr2 = O.hetero_decoder(x2, mode[:1, 0], PD)
    res["clipped_psm_rel_l2"], res["clipped_rm_rel_l2"] = rel_l2(psm2.cpu(), rp2), rel_l2(rm2.cpu(), rr2)
    assert res["clipped_psm_rel_l2"] < 1e-3 and res["clipped_rm_rel_l2"] < 1e-3, res
    # heights that are multiples of 8 but not of the 16-row tile: the second 8-row half of the last tile row lies outside
    # the map (TMA zero fill on the way in, no store on the way out); incl. a map of a single half tile
    for (h3, w3) in ((24, 40), (8, 8), (40, 16)):
        x3 = torch.randn(2, 256, h3, w3)
        with torch.no_grad():
            psm3, rm3 = dec(x3.to(DEV), mode[:2].to(DEV), use_upsample=False)
        rp3, rr3 = O.hetero_decoder(x3, mode[:2, 0], PD)
        e3 = max(rel_l2(psm3.cpu(), rp3), rel_l2(rm3.cpu(), rr3))
        res[f"half_tile_{h3}x{w3}_rel_l2"] = e3
        assert e3 < 1e-3, res
    # anchor_number != 2: the heads run as their own launch (det_heads_kernel) instead of inside the last layer's epilogue
    PD3 = O.synth_decoder_state_dict(13, anchor_number=3)
    dec3 = pkg().HeteroDecoder({"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 3}).eval()
    dec3.load_state_dict(PD3, strict=True)
    dec3 = dec3.to(DEV)
    x4 = torch.randn(2, 256, 16, 24)
    with torch.no_grad():
        psm4, rm4 = dec3(x4.to(DEV), mode[:2].to(DEV), use_upsample=False)
    rp4, rr4 = O.hetero_decoder(x4, mode[:2, 0], PD3)
    assert tuple(psm4.shape) == (2, 3, 16, 24) and tuple(rm4.shape) == (2, 21, 16, 24)
    res["anchor3_psm_rel_l2"], res["anchor3_rm_rel_l2"] = rel_l2(psm4.cpu(), rp4), rel_l2(rm4.cpu(), rr4)
    assert res["anchor3_psm_rel_l2"] < 1e-3 and res["anchor3_rm_rel_l2"] < 1e-3, res
    return res


def check_decoder_logits_golden():
    """psm / rm produced END TO END on the GPU (hmvit_fusion_forward -> hmvit_decoder_forward) against the reference
    pipeline's logits (tests/golden/logits_c256.npz: HeteroFusion -> HeteroDecoder of the unmodified reference).
    Stated tolerance: rel-L2 <= 1e-3 per tensor (north_star: detection-head logits)."""
    g = np.load(os.path.join(GOLDEN, "fusion_c256.npz"))
    gl = np.load(os.path.join(GOLDEN, "logits_c256.npz"))
    C, B, L, H, W, seed = (int(v) for v in g["meta"][:6])
    cfg = O.default_config(input_dim=C)
    P = O.synth_state_dict(cfg, seed)
    PD = O.synth_decoder_state_dict(seed + 1)
    x, T, mode, rl, mask = O.synth_inputs(B, L, C, H, W, g["record_len"].tolist(), seed + 100,
                                          tx=float(g["meta"][6]), ty=float(g["meta"][7]))
    net = pkg().HeteroFusion(cfg).eval()
    net.load_state_dict(P, strict=True)
    dec = pkg().HeteroDecoder({"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}).eval()
    dec.load_state_dict(PD, strict=True)
    net, dec = net.to(DEV), dec.to(DEV)
    with torch.no_grad():
        y = net(x.to(DEV), T.to(DEV), mode.to(DEV), rl.to(DEV), mask.to(DEV))
        psm, rm = dec(y.unsqueeze(1), mode.to(DEV), use_upsample=False)
    res = {"psm_rel_l2": rel_l2(psm.cpu(), torch.from_numpy(gl["psm"])), "rm_rel_l2": rel_l2(rm.cpu(), torch.from_numpy(gl["rm"]))}
    assert res["psm_rel_l2"] < 1e-3 and res["rm_rel_l2"] < 1e-3, res
    return res


def check_model_glue():
    """Model glue behind the encoders (bevformer_point_pillar_hetero.py:113-134): per-modality BEV features -> combine ->
    regroup -> fusion -> decoder -> psm / rm on the GPU, against the loop restatement of the same pipeline on the CPU
    oracle.  Mixed modalities, a ragged batch.  Stated tolerance for this COMPOSED case: rel-L2 <= 2e-3 per tensor -- each
    stage has its own 1e-3 bar against the oracle on identical inputs (fusion checks; decoder_vs_oracle), and the logits of
    the reference-pinned case meet 1e-3 end to end (decoder_logits_golden); here the two stages' errors add on unnormalised
    random features (measured 1.5e-3 / 1.0e-3)."""
    from importlib import import_module
    M = import_module("hmvit_b200.model")
    torch.manual_seed(5)
    B, L, H, W = 2, 3, 16, 24
    mode = torch.tensor([[1, 0, 1], [0, 1, 0]], dtype=torch.int32)
    rl = torch.tensor([3, 2])
    cfgf = O.default_config()
    P, PD = O.synth_state_dict(cfgf, 21), O.synth_decoder_state_dict(22)
    _, T, _, _, _ = O.synth_inputs(B, L, 256, H, W, rl.tolist(), 23)
    mu = O.unpad_mode_encoding(mode, rl)
    cam = torch.randn(int((mu == 0).sum()), 256, H, W)
    lid = torch.randn(int((mu == 1).sum()), 256, H, W)
    cfg = {"max_cav": L, "compression": 0, "anchor_number": 2, "spatial_transform": cfgf["spatial_transform"],
           "hetero_fusion": cfgf, "hetero_decoder": {"input_dim": 256, "num_layer": 2, "num_ch_dec": [256, 256], "anchor_number": 2}}
    net = M.BevformerPointPillarHetero(cfg).eval()
    net.fusion_net.load_state_dict(P, strict=True)
    net.decoder.load_state_dict(PD, strict=True)
    net = net.to(DEV)
    with torch.no_grad():
        out = net.forward_features(cam.to(DEV), lid.to(DEV), mode.to(DEV), rl.to(DEV), T.to(DEV))
    rp, rr = O.detector_forward_features(cam, lid, mode, rl, T, P, PD, cfgf)
    res = {"psm_rel_l2": rel_l2(out["psm"].cpu(), rp), "rm_rel_l2": rel_l2(out["rm"].cpu(), rr)}
    assert res["psm_rel_l2"] < 2e-3 and res["rm_rel_l2"] < 2e-3, res
    return res


def check_postprocess_golden():
    """Detection post-processing on the GPU (hmvit_postprocess: threshold, anchor decoding, corners, projection, filters,
    rotated NMS, range mask) against the outputs of the UNMODIFIED reference post-processor (tests/golden/postproc.npz;
    shapely's polygon areas replaced by the oracle's convex clipping there: that part is parity-unpinned).  Bar: the SAME
    boxes in the SAME order (count and score order), corners within 1e-4 m, scores within 1e-6."""
    sys.path.insert(0, GOLDEN)
    import make_golden_postproc as G
    g = np.load(os.path.join(GOLDEN, "postproc.npz"))
    res = {}
    for name, (H, W, seed, density) in G.CASES.items():
        P = G.params(H, W)
        pp = pkg().VoxelPostprocessor(P, train=False)
        anchors = pp.generate_anchor_box()
        psm, rm = G.synth_outputs(H, W, 2, seed, density)
        T = torch.from_numpy(g[f"{name}_T"])
        for tag in ("proj", "noproj"):
            cav = {"transformation_matrix": T, "anchor_box": torch.from_numpy(anchors)}
            if tag == "noproj":
                cav["no_post_projection"] = True
            boxes, scores = pp.post_process({"ego": cav}, {"ego": {"psm": psm.to(DEV), "rm": rm.to(DEV)}})
            rb, rs = torch.from_numpy(g[f"{name}_{tag}_boxes"]), torch.from_numpy(g[f"{name}_{tag}_scores"])
            assert tuple(boxes.shape) == tuple(rb.shape), (name, tag, tuple(boxes.shape), tuple(rb.shape))
            db, ds = float((boxes.cpu() - rb).abs().max()), float((scores.cpu() - rs).abs().max())
            res[f"{name}_{tag}"] = {"boxes": int(rb.shape[0]), "max_abs_corner_diff": db, "max_abs_score_diff": ds}
            assert db < 1e-4 and ds < 1e-6, res
    # nothing above the threshold: (None, None) like the reference (:313-314)
    none = pp.post_process({"ego": cav}, {"ego": {"psm": torch.full((1, 2, 48, 176), -9.0, device=DEV), "rm": rm.to(DEV)}})
    assert none == (None, None)
    return res


def check_fusion_config5_scene():
    """BASELINE config 5 shape, one scene: 7 agents (LiDAR ego + 6 camera collaborators), 256x96x352."""
    cfg, P, inp, y, net = _fusion_case(1, 7, 96, 352, [7], seed=1239, mode=[[1, 0, 0, 0, 0, 0, 0]], tx=100.0, ty=30.0)
    ref = O.hetero_fusion(*inp, P, cfg)
    res = {"rel_l2_vs_oracle": rel_l2(y, ref), "max_rel_vs_oracle": max_rel(y, ref),
           "rel_l2_vs_reference_sample": _vs_reference_sample("config5_scene", y)}
    assert res["rel_l2_vs_oracle"] < 1e-3 and res["rel_l2_vs_reference_sample"] < 1e-3, res
    return res


def _vs_reference_sample(name, y):
    """rel-L2 of the fused feature against the REFERENCE's own output at the BASELINE config shapes (strided sample,
    tests/golden/fusion_configs.npz made by tests/golden/make_golden_configs.py from the same seeded inputs)."""
    g = np.load(os.path.join(GOLDEN, "fusion_configs.npz"))
    sc, sh, sw = (int(v) for v in g[name + "_strides"])
    return rel_l2(y[:, ::sc, ::sh, ::sw], torch.from_numpy(g[name + "_sample"]))


def check_fusion_config1():
    """BASELINE config 1: 2 agents (LiDAR ego + camera collaborator), 256x48x176, batch 1."""
    cfg, P, inp, y, net = _fusion_case(1, 2, 48, 176, [2], seed=1235, mode=[[1, 0]])
    ref = O.hetero_fusion(*inp, P, cfg)
    res = {"rel_l2_vs_oracle": rel_l2(y, ref), "max_rel_vs_oracle": max_rel(y, ref),
           "rel_l2_vs_reference_sample": _vs_reference_sample("config1", y)}
    assert res["rel_l2_vs_oracle"] < 1e-3 and res["rel_l2_vs_reference_sample"] < 1e-3, res
    return res


def check_fusion_config2_scene():
    """BASELINE config 2 shape, one scene (5 mixed agents, 48x176) + ragged second scene."""
    cfg, P, inp, y, net = _fusion_case(2, 5, 48, 176, [5, 3], seed=1236)
    ref = O.hetero_fusion(*inp, P, cfg)
    res = {"rel_l2_vs_oracle": rel_l2(y, ref), "max_rel_vs_oracle": max_rel(y, ref),
           "rel_l2_vs_reference_sample": _vs_reference_sample("config2_scene", y)}
    assert res["rel_l2_vs_oracle"] < 1e-3 and res["rel_l2_vs_reference_sample"] < 1e-3, res
    # exact dead-query elimination: same result without it, up to nothing (identical kernels on slot 0)
    net.skip_dead_queries = False
    x, T, md, rl, mask = inp
    with torch.no_grad():
        y2 = net(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    res["skip_dead_max_abs_diff"] = float((y - y2).abs().max())
    assert res["skip_dead_max_abs_diff"] == 0.0, res
    return res


def check_fusion_properties():
    """Size-independent properties at the bench shape (B=2 to keep the oracle out of the loop):
    (1) scene independence: permuting scenes permutes outputs bit-exactly;
    (2) invisible collaborators: moving a collaborator far outside the map == removing it."""
    cfg, P, net = _mk_module(0)
    x, T, md, rl, mask = _scene(2, 5, 48, 176, [5, 4], seed=77)
    run = lambda *a: net(*[t.to(DEV) for t in a]).cpu()
    with torch.no_grad():
        y = run(x, T, md, rl, mask)
        perm = torch.tensor([1, 0])
        yp = run(x[perm], T[perm], md[perm], rl[perm], mask[perm])
    res = {"scene_perm_max_abs": float((yp - y[perm]).abs().max())}
    assert res["scene_perm_max_abs"] == 0.0, res
    # scene 1 has 4 agents: push agent 3 2 km away (both directions of every pair that involves it)
    T2 = T.clone()
    far = torch.eye(4)
    far[0, 3] = 2000.0
    for j in range(3):
        T2[1, 3, j] = far
        T2[1, j, 3] = far
    rl3, mask3 = rl.clone(), mask.clone()
    rl3[1] = 3
    mask3[1, 3] = 0
    x3 = x.clone()
    x3[1, 3] = 0
    md3 = md.clone()
    md3[1, 3] = 0
    with torch.no_grad():
        ya = run(x, T2, md, rl, mask)
        yb = run(x3, T2, md3, rl3, mask3)
    res["far_agent_equals_removed_rel_l2"] = rel_l2(ya[1], yb[1])
    # equal up to the order of the online-softmax updates (agent 3 contributes exp(-inf) = 0 exactly);
    # the K/V variants computed differ only if the removed agent was the sole one of its type
    assert res["far_agent_equals_removed_rel_l2"] < 1e-6, res
    return res


def check_errors():
    """Error behaviour mirrors the reference: ValueError for unknown architect_mode / bad shapes."""
    import pytest
    p = pkg()
    cfg = O.default_config()
    bad = O.default_config()
    bad["hetero_fusion_block"]["architect_mode"] = "parallel"
    net = p.HeteroFusion(bad).eval().to(DEV)
    x, T, md, rl, mask = _scene(1, 2, 16, 16, [2], seed=1)
    with torch.no_grad():
        with pytest.raises(ValueError):
            net(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV))
        net2 = p.HeteroFusion(cfg).eval().to(DEV)
        with pytest.raises(ValueError):
            net2(x[..., :12, :].contiguous().to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV))
        with pytest.raises(ValueError):
            net2(x, T, md, rl, mask)          # CPU tensors: no fallback
        with pytest.raises(ValueError):
            net2(x.to(DEV), T[:, :1].contiguous().to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV))   # pairwise_t_matrix shape
    bad_bias = p.HeteroFusion(cfg).eval().to(DEV)
    with torch.no_grad():
        bad_bias.hetero_fusion_block.window_attention.relative_position_bias_table.weight.fill_(100.0)
        with pytest.raises(ValueError):
            bad_bias(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV))                   # bias table beyond +-40
    return {"ok": 1}


CHECKS = {
    "gemm_qkv": lambda: check_rowgemm("QKV"),
    "gemm_qkv_ego": lambda: check_rowgemm("QKV_EGO"),
    "gemm_qkv_noln": lambda: check_rowgemm("QKV_NOLN"),
    "gemm_out": lambda: check_rowgemm("OUT"),
    "gemm_out_ego": lambda: check_rowgemm("OUT_EGO"),
    "gemm_ffn1": lambda: check_rowgemm("FFN1"),
    "gemm_ffn2": lambda: check_rowgemm("FFN2"),
    "gemm_head1": lambda: check_rowgemm("HEAD1"),
    "gemm_head2": lambda: check_rowgemm("HEAD2"),
    "gemm_qkv_n8448": lambda: check_rowgemm("QKV", N=8448 // 8),
    "chain": check_chain,
    "chain_ego_n1056": lambda: check_chain(N=1056, ego_only=True),
    "warp_mask_golden": check_warp_mask_golden,
    "mask_adversarial": check_mask_adversarial,
    "attention_golden": check_attention_golden,
    "fusion_small": check_fusion_small,
    "fusion_golden": check_fusion_golden,
    "fusion_repeatable": check_fusion_repeatable,
    "logits_golden": check_logits_golden,
    "decoder_vs_oracle": check_decoder_vs_oracle,
    "decoder_logits_golden": check_decoder_logits_golden,
    "model_glue": check_model_glue,
    "postprocess_golden": check_postprocess_golden,
    "fusion_config1": check_fusion_config1,
    "fusion_config2_scene": check_fusion_config2_scene,
    "fusion_properties": check_fusion_properties,
    "fusion_config5_scene": check_fusion_config5_scene,
    "errors": check_errors,
}


# ----------------------------------------------------------------------------------------------
def check_cuda_graph():
    """The whole forward only enqueues work on the caller's stream (no host sync, no allocation inside the
    C-ABI): it can be captured into a CUDA graph and replayed; the replay is bit-identical."""
    cfg, P, net = _mk_module(0)
    x, T, md, rl, mask = _scene(2, 4, 32, 48, [4, 3], seed=9)
    inp = [t.to(DEV) for t in (x, T, md, rl, mask)]
    with torch.no_grad():
        y_ref = net(*inp).clone()
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            net(*inp)                                  # warm-up on the capture stream (packs weights, sizes workspace)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                y_g = net(*inp)
        torch.cuda.current_stream().wait_stream(s)
        inp[0].mul_(0.5)                               # new input values in the captured buffer
        y_new_ref = net(*inp).clone()
        g.replay()
        torch.cuda.synchronize()
    res = {"replay_vs_eager_max_abs": float((y_g - y_new_ref).abs().max()), "changed": float((y_new_ref - y_ref).abs().max())}
    assert res["replay_vs_eager_max_abs"] == 0.0 and res["changed"] > 0.0, res
    return res


def check_ragged_batch():
    """Edge cases of the batch structure: single-agent scenes (record_len 1), fully populated scenes and
    mixed record_len in one batch, all-camera and all-LiDAR scenes; against the oracle."""
    cfg, P, net = _mk_module(1)
    B, L, H, W = 4, 5, 16, 24
    mode = [[1, 0, 0, 0, 0], [0, 0, 0, 0, 0], [1, 1, 1, 1, 1], [0, 1, 0, 1, 0]]
    x, T, md, rl, mask = _scene(B, L, H, W, [1, 5, 3, 2], seed=21, mode=mode, tx=12, ty=6)
    with torch.no_grad():
        y = net(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    ref = O.hetero_fusion(x, T, md, rl, mask, P, cfg)
    res = {"rel_l2_vs_oracle": rel_l2(y, ref), "per_scene": [round(rel_l2(y[b], ref[b]), 6) for b in range(B)]}
    assert all(v < 1e-3 for v in res["per_scene"]), res
    return res


def check_partial_tile():
    """A map whose token count is not a multiple of the 128-token GEMM tile (24 x 24 = 576 = 4.5 tiles): the last tile
    of every agent is half empty in the QKV / chain / head kernels (clamped loads, clipped stores); against the oracle."""
    cfg, P, net = _mk_module(1)
    B, L, H, W = 2, 3, 24, 24
    x, T, md, rl, mask = _scene(B, L, H, W, [3, 2], seed=33, mode=[[1, 0, 1], [0, 1, 0]], tx=8, ty=8)
    with torch.no_grad():
        y = net(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
        yb = net.hetero_fusion_block(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    ref = O.hetero_fusion(x, T, md, rl, mask, P, cfg)
    refb = O.hetero_fusion_block(x, T, md, rl, mask, P, cfg["hetero_fusion_block"]) if hasattr(O, "hetero_fusion_block") else None
    res = {"rel_l2_vs_oracle": rel_l2(y, ref), "finite": bool(torch.isfinite(y).all())}
    if refb is not None:
        errs = [rel_l2(yb[b, l], refb[b, l]) for b in range(B) for l in range(int(rl[b]))]
        res["block_worst_rel_l2"] = max(errs)
        assert res["block_worst_rel_l2"] < 1e-3, res
    assert res["finite"] and res["rel_l2_vs_oracle"] < 1e-3, res
    return res


CHECKS.update({"cuda_graph": check_cuda_graph, "ragged_batch": check_ragged_batch, "partial_tile": check_partial_tile})
# (attn_split_vs_single is registered next to the backward checks below: it uses tests/emul_ops.py)


# ----------------------------------------------------------------------------------------------
# backward / training path: every CUDA kernel against tests/emul_ops.py, then whole-module
# gradients against torch.autograd through the CPU oracle
# ----------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emul_ops as EM  # noqa: E402


def _geo_small(seed=7, B=2, L=3, H=16, W=24, record_len=(3, 2)):
    x, T, md, rl, mask = _scene(B, L, H, W, list(record_len), seed=seed, tx=10, ty=5)
    return dict(B=B, L=L, H=H, W=W, N=H * W, T=T, mode=md.to(torch.int32), rl=rl.to(torch.int32), cav=mask.to(torch.int32))


def _valid_rows(g, t, rows=False, ego_only=False):
    """select the data of active agents from a cm (B*L, 256, N) or rows (R, 256) tensor"""
    out = []
    for b, l, ai in EM._agents(g["B"], g["L"], g["rl"], ego_only):
        out.append(t[ai * g["N"]:(ai + 1) * g["N"]] if rows else t[ai])
    return torch.stack(out).float()


def check_bwd_small_kernels():
    """row statistics, LayerNorm backward, GELU backward, bf16 cast, bias column sums vs the torch emulation."""
    ops = pkg().ops
    g = _geo_small()
    B, L, N = g["B"], g["L"], g["N"]
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B * L, 256, N, generator=gen) * 1.5 + 0.3
    dz = torch.randn(B * L, 256, N, generator=gen)
    dres = torch.randn(B * L, 256, N, generator=gen)
    md, rl = g["mode"].to(DEV), g["rl"].to(DEV)
    res = {}
    for ego in (False, True):
        st_ref = torch.zeros(B * L * N, 2)
        EM.bwd_row_stats(x, st_ref, B=B, L=L, N=N, record_len=g["rl"], ego_only=ego)
        st = torch.zeros(B * L * N, 2, device=DEV)
        ops.bwd_row_stats(x.to(DEV), st, B=B, L=L, N=N, record_len=rl, ego_only=ego)
        res[f"stats_ego{int(ego)}"] = max_rel(st.cpu(), st_ref)
        dx_ref = torch.zeros_like(x)
        EM.bwd_layernorm(dz, x, st_ref, dres, dx_ref, B=B, L=L, N=N, record_len=g["rl"], ego_only=ego)
        dx = torch.zeros_like(x, device=DEV)
        ops.bwd_layernorm(dz.to(DEV), x.to(DEV), st, dres.to(DEV), dx, B=B, L=L, N=N, record_len=rl, ego_only=ego)
        res[f"ln_bwd_ego{int(ego)}"] = rel_l2(dx.cpu(), dx_ref)
        for rows in (False, True):
            y = (torch.randn(B * L * N, 256, generator=gen).to(torch.bfloat16) if rows else dz)
            db_ref = torch.zeros(2, 1280)
            EM.bwd_colsum(y, db_ref[:, 256:], B=B, L=L, N=N, mode=g["mode"], record_len=g["rl"], ego_only=ego)
            db = torch.zeros(2, 1280, device=DEV)
            ops.bwd_colsum(y.to(DEV), db[:, 256:], B=B, L=L, N=N, mode=md, record_len=rl, ego_only=ego)
            res[f"colsum_rows{int(rows)}_ego{int(ego)}"] = max_rel(db.cpu(), db_ref)
    hp, dh = torch.randn(B * L, 256, N, generator=gen) * 2, torch.randn(B * L, 256, N, generator=gen)
    hp_d, dh_d = hp.to(DEV), dh.to(DEV)
    EM.bwd_gelu(hp, dh)
    ops.bwd_gelu(hp_d, dh_d)
    res["gelu_fwd"], res["gelu_bwd"] = max_rel(hp_d.cpu(), hp), max_rel(dh_d.cpu(), dh)
    src = torch.randn(5, 1000, 256, generator=gen)
    dst = torch.empty(5, 1000, 256, dtype=torch.bfloat16, device=DEV)
    ops.bwd_cast_bf16(src.to(DEV), dst)
    assert torch.equal(dst.cpu(), src.to(torch.bfloat16))
    torch.cuda.synchronize()
    assert all(v < 2e-5 for v in res.values()), res
    return res


def check_bwd_wgrad():
    """typed weight gradient, all operand-layout combinations (+ on-the-fly normalisation): cm x cm, cm x cm + LN and rows x cm run the
    tcgen05 kernel (csrc/wgrad_tc.cuh), cm x rows the same kernel with the operands exchanged, rows x rows the wmma kernel (csrc/bwd.cuh);
    operands are rounded to bf16 while staged, fp32 accumulate."""
    ops = pkg().ops
    g = _geo_small(seed=8)
    B, L, N = g["B"], g["L"], g["N"]
    gen = torch.Generator().manual_seed(2)
    md, rl = g["mode"].to(DEV), g["rl"].to(DEV)
    a_cm, b_cm = torch.randn(B * L, 256, N, generator=gen), torch.randn(B * L, 256, N, generator=gen) + 0.5
    a_rows = torch.randn(B * L * N, 256, generator=gen).to(torch.bfloat16)
    b_rows = torch.randn(B * L * N, 256, generator=gen).to(torch.bfloat16)
    st = torch.zeros(B * L * N, 2)
    EM.bwd_row_stats(b_cm, st, B=B, L=L, N=N, record_len=g["rl"])
    res = {}
    for name, a, b, stats, ego in (("cm_cm", a_cm, b_cm, None, False), ("cm_cmln", a_cm, b_cm, st, False),
                                   ("cm_rows", a_cm, b_rows, None, True), ("rows_cm", a_rows, b_cm, st, False),
                                   ("rows_rows", a_rows, b_rows, None, False)):
        ref = torch.zeros(2, 1280, 256)
        EM.bwd_wgrad(a, b, ref, B=B, L=L, N=N, mode=g["mode"], record_len=g["rl"], ego_only=ego, b_stats=stats, row0=512)
        dw = torch.zeros(2, 1280, 256, device=DEV)
        ops.bwd_wgrad(a.to(DEV), b.to(DEV), dw, B=B, L=L, N=N, mode=md, record_len=rl, ego_only=ego,
                      b_stats=None if stats is None else stats.to(DEV), row0=512)
        res[name] = rel_l2(dw.cpu(), ref)
        assert float(dw[:, :512].abs().max()) == 0.0 and float(dw[:, 768:].abs().max()) == 0.0
    torch.cuda.synchronize()
    assert all(v < 5e-3 for v in res.values()), res          # operands rounded to bf16 while staged, fp32 accumulate
    return res


def check_bwd_dgrad_cat():
    """input gradient of the fused Q | K' | V' projection as one K = 1280 GEMM (csrc/dgrad_cat.cuh) against the emulation on the
    same bf16 operands; 16 x 24 (whole 128-token tiles) and 24 x 24 (576 tokens: a half tile at the end of every agent)."""
    ops = pkg().ops
    res = {}
    for name, kw in (("16x24", {}), ("24x24", dict(H=24, W=24))):
        g = _geo_small(seed=11, **kw)
        B, L, N = g["B"], g["L"], g["N"]
        gen = torch.Generator().manual_seed(5)
        dcat = torch.randn(5, B * L * N, 256, generator=gen).to(torch.bfloat16)
        w = [(torch.randn(1280, 256, generator=gen) / 16).to(torch.bfloat16) for _ in range(2)]
        ref = torch.full((B * L, 256, N), 7.0)
        EM.bwd_dgrad_cat(dcat, w[0], w[1], ref, B=B, L=L, N=N, mode=g["mode"], record_len=g["rl"])
        out = torch.full((B * L, 256, N), 7.0, device=DEV)
        ops.bwd_dgrad_cat(dcat.to(DEV), w[0].to(DEV), w[1].to(DEV), out, B=B, L=L, N=N, mode=g["mode"].to(DEV), record_len=g["rl"].to(DEV))
        res[name] = rel_l2(out.cpu(), ref)
        pad = [ai for ai in range(B * L) if ai % L >= int(g["rl"][ai // L])]
        assert pad and all(bool((out[ai] == 7.0).all()) for ai in pad), "padded slots must stay untouched"
    torch.cuda.synchronize()
    assert all(v < 1e-5 for v in res.values()), res          # same bf16 operands, fp32 accumulate: summation order only
    return res


def check_bwd_cast_colsum():
    """bf16 copy of the five gradient planes + their typed column sums in one pass (cast_colsum_kernel) vs the emulation: the copy
    bit-exact, the sums to fp32 summation order (random data in every slot, so padded slots count on both sides)."""
    ops = pkg().ops
    g = _geo_small(seed=12, H=24, W=24)
    B, L, N = g["B"], g["L"], g["N"]
    gen = torch.Generator().manual_seed(6)
    src = torch.randn(5, B * L * N, 256, generator=gen)
    ref_dst, ref_db = torch.empty(5, B * L * N, 256, dtype=torch.bfloat16), torch.full((2, 1300), 0.25)
    EM.bwd_cast_colsum(src, ref_dst, ref_db, B=B, L=L, N=N, mode=g["mode"])
    dst, db = torch.empty(5, B * L * N, 256, dtype=torch.bfloat16, device=DEV), torch.full((2, 1300), 0.25, device=DEV)
    ops.bwd_cast_colsum(src.to(DEV), dst, db, B=B, L=L, N=N, mode=g["mode"].to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), ref_dst)
    assert bool((db[:, 1280:] == 0.25).all()), "columns past the 5 x 256 sums must stay untouched"
    res = {"colsum_max_rel": max_rel(db.cpu()[:, :1280], ref_db[:, :1280])}
    assert res["colsum_max_rel"] < 2e-5, res
    return res


def check_bwd_lin_variants():
    """the row-GEMM variants the backward uses (7-10) against the emulation with equally rounded operands."""
    p = pkg()
    ops, lib = p.ops, p._lib
    g = _geo_small(seed=9)
    B, L, N = g["B"], g["L"], g["N"]
    gen = torch.Generator().manual_seed(3)
    md, rl = g["mode"].to(DEV), g["rl"].to(DEV)
    x = torch.randn(B * L, 256, N, generator=gen)
    rows = torch.randn(B * L * N, 256, generator=gen).to(torch.bfloat16)
    w = [torch.randn(256, 256, generator=gen) / 16 for _ in range(2)]
    bias = torch.randn(2, 256, generator=gen) * 0.1
    resid = torch.randn(B * L, 256, N, generator=gen)
    res = {}
    common = dict(B=B, L=L, N=N, n_out=256)
    for name, variant, a, wd, kw, ego in (
            ("ln_lin_cm", lib.GEMM_LN_LIN_CM, x, [tf32(t) for t in w], {}, False),
            ("lin_cm", lib.GEMM_LIN_CM, x, [tf32(t) for t in w], {}, True),
            ("lin_rows", lib.GEMM_LIN_ROWS, x, [tf32(t) for t in w], {}, False),
            ("rows_lin_cm", lib.GEMM_ROWS_LIN_CM, rows, [t.to(torch.bfloat16) for t in w], {}, False),
            ("rows_lin_cm_resid", lib.GEMM_ROWS_LIN_CM, rows, [t.to(torch.bfloat16) for t in w], {"resid": resid}, True)):
        to_rows = variant == lib.GEMM_LIN_ROWS
        ref = torch.zeros(B * L * N, 256) if to_rows else torch.zeros(B * L, 256, N)
        EM.rowgemm(variant, a=a.float(), w0=wd[0].float(), w1=wd[1].float(), bias=bias, out=ref, mode=g["mode"], record_len=g["rl"],
                   ego_only=ego, **kw, **common)
        out = torch.zeros(B * L * N, 256, dtype=torch.bfloat16, device=DEV) if to_rows else torch.zeros(B * L, 256, N, device=DEV)
        ops.rowgemm(variant, a=a.to(DEV), w0=wd[0].to(DEV), w1=wd[1].to(DEV), bias=bias.to(DEV), out=out, mode=md, record_len=rl,
                    ego_only=ego, **{k: v.to(DEV) for k, v in kw.items()}, **common)
        res[name] = rel_l2(out.float().cpu(), ref)
    torch.cuda.synchronize()
    assert all(v < 4e-3 for v in res.values()), res          # tf32 / bf16 operand rounding of A inside the kernel
    return res


def _attn_case(kind, ego_only, seed):
    p = pkg()
    ops = p.ops
    g = _geo_small(seed=seed)
    B, L, H, W, N = g["B"], g["L"], g["H"], g["W"], g["N"]
    R = B * L * N
    gen = torch.Generator().manual_seed(seed)
    q = (torch.randn(R, 256, generator=gen) * 0.6).to(torch.bfloat16)
    k = (torch.randn(2, R, 256, generator=gen) * 0.6).to(torch.bfloat16)
    v = torch.randn(2, R, 256, generator=gen).to(torch.bfloat16)
    bk, bv = torch.randn(2, 2, 256, generator=gen) * 0.1, torch.randn(2, 2, 256, generator=gen) * 0.1
    table = torch.randn(225, 8, generator=gen)
    d_o = torch.randn(R, 256, generator=gen).to(torch.bfloat16)
    geo = dict(B=B, L=L, H=H, W=W, kind=kind, cell=1.6, ego_only=ego_only)
    cpu = dict(mode=g["mode"], record_len=g["rl"], cav_mask=g["cav"], T=g["T"])
    dev = {kk: vv.to(DEV).contiguous() for kk, vv in cpu.items()}
    return ops, g, geo, cpu, dev, (q, k, v, bk, bv, table, d_o)


def check_attn_bwd():
    """attention forward statistics (lse) and the attention backward kernel against autograd of the emulated
    attention (window and grid partitions, all egos and ego-only)."""
    res = {}
    for kind, ego_only, seed in ((0, False, 31), (1, False, 32), (1, True, 33)):
        ops, g, geo, cpu, dev, (q, k, v, bk, bv, table, d_o) = _attn_case(kind, ego_only, seed)
        R = g["B"] * g["L"] * g["N"]
        out = torch.zeros(R, 256, dtype=torch.bfloat16, device=DEV)
        lse = torch.zeros(R, 8, device=DEV)
        ops.group_attn(q=q.to(DEV), k=k.to(DEV), v=v.to(DEV), bk=bk.to(DEV), bv=bv.to(DEV), bias_table=table.to(DEV), out=out,
                       lse=lse, **geo, **dev)
        out_ref, lse_ref = torch.zeros(R, 256), torch.zeros(R, 8)
        EM.group_attn(q=q.float(), k=k.float(), v=v.float(), bk=bk, bv=bv, bias_table=table, out=out_ref, lse=lse_ref, **geo, **cpu)
        tag = f"kind{kind}_ego{int(ego_only)}"
        res[f"out_{tag}"] = rel_l2(out.float().cpu(), out_ref)
        res[f"lse_abs_{tag}"] = float((lse.cpu() - lse_ref).abs().max())
        gd = {n: torch.zeros(s, device=DEV) for n, s in (("dq", (R, 256)), ("dk", (2, R, 256)), ("dv", (2, R, 256)),
                                                          ("dbk", (2, 2, 256)), ("dbv", (2, 2, 256)), ("dbias_table", (225, 8)))}
        ops.group_attn_bwd(q=q.to(DEV), k=k.to(DEV), v=v.to(DEV), bk=bk.to(DEV), bv=bv.to(DEV), bias_table=table.to(DEV),
                           o=out, d_o=d_o.to(DEV), lse=lse, **gd, **geo, **dev)
        gr = {n: torch.zeros_like(t, device="cpu") for n, t in gd.items()}
        EM.group_attn_bwd(q=q.float(), k=k.float(), v=v.float(), bk=bk, bv=bv, bias_table=table, o=out_ref, d_o=d_o.float(),
                          lse=lse_ref, **gr, **geo, **cpu)
        for n in gd:
            res[f"{n}_{tag}"] = rel_l2(gd[n].cpu(), gr[n])
    torch.cuda.synchronize()
    for n, val in res.items():
        tol = 0.03 if n.startswith("lse_abs") else 2e-2      # bf16 tap blend, bf16 P / dS operands, fp32 accumulate
        assert val < tol, (n, val, res)
    return res


def check_attn_split_vs_single():
    """The three implementations of hmvit_group_attn (fused persistent tcgen05 kernel = default, split, single) against the
    emulated attention: 7 agents (two tap passes in the single kernel), ragged record_len, window / grid / ego-only,
    softmax statistics; plus a NON-identity diagonal transform (the ego's own keys then take the general warp path)."""
    res = {}
    for kind, ego_only, seed, twist in ((0, False, 41, False), (1, False, 42, False), (1, True, 43, False), (1, False, 44, True)):
        p = pkg()
        ops = p.ops
        g = _geo_small(seed=seed, B=2, L=7, H=16, W=24, record_len=(7, 4))
        if twist:
            g["T"][0, 1, 1] = g["T"][0, 2, 1].clone()          # slot 1 of scene 0 warps its own features
        B, L, H, W, N = g["B"], g["L"], g["H"], g["W"], g["N"]
        R = B * L * N
        gen = torch.Generator().manual_seed(seed)
        q = (torch.randn(R, 256, generator=gen) * 0.6).to(torch.bfloat16)
        k = (torch.randn(2, R, 256, generator=gen) * 0.6).to(torch.bfloat16)
        v = torch.randn(2, R, 256, generator=gen).to(torch.bfloat16)
        bk, bv = torch.randn(2, 2, 256, generator=gen) * 0.1, torch.randn(2, 2, 256, generator=gen) * 0.1
        table = torch.randn(225, 8, generator=gen)
        geo = dict(B=B, L=L, H=H, W=W, kind=kind, cell=1.6, ego_only=ego_only)
        cpu = dict(mode=g["mode"], record_len=g["rl"], cav_mask=g["cav"], T=g["T"])
        dev = {kk: vv.to(DEV).contiguous() for kk, vv in cpu.items()}
        outs, lses = {}, {}
        for impl in ("fused", "split", "single"):
            outs[impl] = torch.zeros(R, 256, dtype=torch.bfloat16, device=DEV)
            lses[impl] = torch.zeros(R, 8, device=DEV)
            ops.group_attn(q=q.to(DEV), k=k.to(DEV), v=v.to(DEV), bk=bk.to(DEV), bv=bv.to(DEV), bias_table=table.to(DEV),
                           out=outs[impl], lse=lses[impl], impl=impl, **geo, **dev)
        out_ref, lse_ref = torch.zeros(R, 256), torch.zeros(R, 8)
        EM.group_attn(q=q.float(), k=k.float(), v=v.float(), bk=bk, bv=bv, bias_table=table, out=out_ref, lse=lse_ref, **geo, **cpu)
        tag = f"kind{kind}_ego{int(ego_only)}" + ("_twist" if twist else "")
        # a query that sees no key at all (possible only with the twisted diagonal) is NaN in the emulation (softmax of an
        # empty set) and 0 in the kernels: compare the rows that have keys
        ok = torch.isfinite(out_ref).all(-1)
        fin = torch.isfinite(lse_ref)
        for impl in ("fused", "split", "single"):
            res[f"{impl}_{tag}"] = rel_l2(outs[impl].float().cpu()[ok], out_ref[ok])
        for impl in ("fused", "split"):
            assert float(outs[impl].float().cpu()[~ok].abs().sum()) == 0.0, (impl, tag)
            assert bool((torch.isfinite(lses[impl].cpu()) == fin).all()), (impl, tag)
            res[f"lse_abs_{impl}_{tag}"] = float((lses[impl].cpu()[fin] - lse_ref[fin]).abs().max())
    torch.cuda.synchronize()
    for n, val in res.items():
        assert val < (0.03 if n.startswith("lse_abs") else 2e-2), (n, val, res)
    return res


def _train_case(B, L, H, W, record_len, seed, mode=None, skip_dead=True):
    cfg = O.default_config()
    cfg["hetero_fusion_block"]["drop_out"] = 0.0
    P = O.synth_state_dict(cfg, 0)
    net = pkg().HeteroFusion(cfg)
    net.load_state_dict(P, strict=True)
    net = net.to(DEV).train()
    net.skip_dead_queries = skip_dead
    x, T, md, rl, mask = _scene(B, L, H, W, record_len, seed, mode=mode, tx=10, ty=5)
    g_out = torch.randn(B, 256, H, W, generator=torch.Generator().manual_seed(seed + 1))
    # reference: autograd through the CPU oracle
    Pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    xr = x.clone().requires_grad_(True)
    y_ref = O.hetero_fusion(xr, T, md, rl, mask, Pg, cfg)
    names = [k for k, v in Pg.items() if v.is_floating_point()]
    gs = torch.autograd.grad((y_ref * g_out).sum(), [xr] + [Pg[k] for k in names], allow_unused=True)
    g_ref = dict(zip(["x"] + names, gs))
    # CUDA path
    xd = x.to(DEV).requires_grad_(True)
    y = net(xd, T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV))
    (y * g_out.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    res = {"fwd_rel_l2": rel_l2(y.detach().cpu(), y_ref.detach()), "dx_rel_l2": rel_l2(xd.grad.cpu(), g_ref["x"])}
    worst, worst_name, n_checked = 0.0, "", 0
    gnorm = max(float(v.norm()) for k, v in g_ref.items() if v is not None and k != "x")
    for name, prm in net.named_parameters():
        ref = g_ref.get(name)
        if ref is None or "aggregate_fc" in name:
            assert prm.grad is None or float(prm.grad.abs().max()) == 0.0, name
            continue
        if float(ref.norm()) < 1e-4 * gnorm:                 # absent modality / negligible: absolute check
            assert prm.grad is None or float((prm.grad.cpu() - ref).norm()) < 1e-3 * gnorm, name
            continue
        e = rel_l2(prm.grad.cpu(), ref)
        n_checked += 1
        if e > worst:
            worst, worst_name = e, name
    res.update({"param_worst_rel_l2": worst, "param_worst": worst_name, "params_checked": n_checked})
    return res


def check_train_grads_small():
    """Whole-module gradients (dL/dx and dL/dtheta of every used parameter) of the CUDA training path against
    torch.autograd through the fp32 CPU oracle.  Stated tolerance (bf16 / tf32 operands, fp32 accumulate):
    rel-L2 <= 1.5e-2 per parameter tensor and <= 5e-3 for dL/dx (measured 5.7e-3 / 8.3e-4; the bar was 3e-2 until the
    end of round 2); the forward output keeps the inference tolerance 1e-3."""
    res = {}
    for tag, args in (("mixed", dict(B=2, L=3, H=16, W=24, record_len=[3, 2], seed=41)),
                      ("nodead", dict(B=1, L=3, H=16, W=24, record_len=[3], seed=42, skip_dead=False)),
                      ("lidar_ego", dict(B=1, L=4, H=32, W=48, record_len=[4], seed=43, mode=[[1, 0, 0, 1]]))):
        r = _train_case(**args)
        res.update({f"{tag}_{k}": v for k, v in r.items()})
        assert r["fwd_rel_l2"] < 1e-3 and r["dx_rel_l2"] < 5e-3 and r["param_worst_rel_l2"] < 1.5e-2 and r["params_checked"] > 30, res
    return res


def check_train_api():
    """grad-mode dispatch of the module surface: train mode with the yaml's drop_out = 0.1 runs (Philox dropout), eval +
    grad works, no_grad is the inference path, the block alone is differentiable, parameters without requires_grad get
    no gradient."""
    cfg = O.default_config()
    P = O.synth_state_dict(cfg, 0)
    net = pkg().HeteroFusion(cfg)
    net.load_state_dict(P, strict=True)
    net = net.to(DEV)
    x, T, md, rl, mask = _scene(1, 2, 16, 16, [2], 3, tx=5, ty=5)
    args = [t.to(DEV) for t in (T, md, rl, mask)]
    net.train()
    yt = net(x.to(DEV), *args)                               # the shipped yaml's drop_out = 0.1: no longer rejected
    assert yt.requires_grad and torch.isfinite(yt).all()
    assert net.hetero_fusion_block.last_dropout_seed is not None
    net.eval()
    with torch.no_grad():
        y0 = net(x.to(DEV), *args)
    y1 = net(x.to(DEV), *args)
    assert y1.requires_grad
    d = rel_l2(y1.detach().cpu(), y0.cpu())
    for prm in net.mlp_head.parameters():
        prm.requires_grad_(False)
    y1.sum().backward()
    assert all(prm.grad is None for prm in net.mlp_head.parameters())
    assert net.hetero_fusion_block.window_attention.relation_att.grad is not None
    xb = x.to(DEV).requires_grad_(True)
    yb = net.hetero_fusion_block(xb, *args)
    yb.square().mean().backward()
    assert xb.grad is not None and torch.isfinite(xb.grad).all()
    assert d < 1e-3, d
    return {"train_vs_infer_fwd_rel_l2": d}


CHECKS.update({"bwd_small_kernels": check_bwd_small_kernels, "bwd_wgrad": check_bwd_wgrad,
               "bwd_lin_variants": check_bwd_lin_variants, "bwd_dgrad_cat": check_bwd_dgrad_cat, "bwd_cast_colsum": check_bwd_cast_colsum, "attn_bwd": check_attn_bwd,
               "train_grads_small": check_train_grads_small, "train_api": check_train_api,
               "attn_split_vs_single": check_attn_split_vs_single})


# ----------------------------------------------------------------------------------------------
# round 2: reference-default initialisation, the shipped yaml's 128 x 128 grid, seed sweep, index probe
# ----------------------------------------------------------------------------------------------
def _default_init_module():
    """The reference's DEFAULT initialisation under torch.manual_seed(0): the product module mirrors the
    reference's registration order, so the same seed gives the same state_dict (asserted against the reference in
    tests/golden/make_golden_r2.py; pinned here by the parameter checksum stored with the goldens)."""
    cfg = O.default_config()
    torch.manual_seed(0)
    net = pkg().HeteroFusion(cfg).eval()
    P = {k: v.clone() for k, v in net.state_dict().items()}
    g = np.load(os.path.join(GOLDEN, "fusion_r2.npz"))
    cs = float(sum(v.double().abs().sum() for v in P.values()))
    assert abs(cs - float(g["default_param_checksum"][0])) <= 1e-9 * cs, "default initialisation drifted from the golden's"
    return cfg, P, net.to(DEV)


R2_CASES = {
    # name: (weights, B, L, record_len, seed, mode, H, W, synth_inputs kwargs)
    "c1_default": ("default", 1, 2, [2], 1235, [[1, 0]], 48, 176, {}),
    "c2_default": ("default", 2, 5, [5, 3], 1236, None, 48, 176, {}),
    "y128_synth": ("synth", 2, 4, [4, 2], 1240, None, 128, 128, {"tx": 60.0, "ty": 60.0}),
    "y128_default": ("default", 2, 4, [4, 2], 1240, None, 128, 128, {"tx": 60.0, "ty": 60.0}),
}


def check_fusion_r2_goldens():
    """Whole forward against the REFERENCE's own outputs (strided sample + norms, tests/golden/fusion_r2.npz) for the two
    round-2 families: reference-default initialisation under torch.manual_seed(0) at the BASELINE config 1 / 2 shapes,
    and the shipped yaml's 256 x 128 x 128 grid with both weight families.  Bar: rel-L2 <= 1e-3 per tensor."""
    g = np.load(os.path.join(GOLDEN, "fusion_r2.npz"))
    res = {}
    for name, (wts, B, L, rl, seed, mode, H, W, kw) in R2_CASES.items():
        cfg, P, net = _default_init_module() if wts == "default" else _mk_module(0)
        x, T, md, rlt, mask = _scene(B, L, H, W, rl, seed, mode=mode, **kw)
        with torch.no_grad():
            y = net(x.to(DEV), T.to(DEV), md.to(DEV), rlt.to(DEV), mask.to(DEV)).cpu()
        sc, sh, sw = (int(v) for v in g[name + "_strides"])
        ref = torch.from_numpy(g[name + "_sample"])
        res[name + "_rel_l2"] = rel_l2(y[:, ::sc, ::sh, ::sw], ref)
        res[name + "_max_rel"] = max_rel(y[:, ::sc, ::sh, ::sw], ref)
        res[name + "_norm_rel"] = abs(float(y.double().norm()) / float(g[name + "_norms"][0]) - 1.0)
        assert res[name + "_rel_l2"] < 1e-3 and res[name + "_norm_rel"] < 1e-3, res
    return res


def check_seed_sweep():
    """Five weight / input seeds (one of them at the full 48 x 176 map) against the fp32 oracle: the worst rel-L2 must stay
    below the 1e-3 bar; max-rel is reported beside it."""
    res, worst = {}, 0.0
    cases = [(s, 2, 4, 32, 48, [4, 3]) for s in (11, 12, 13, 14)] + [(15, 1, 5, 48, 176, [5])]
    for s, B, L, H, W, rl in cases:
        cfg, P, inp, y, net = _fusion_case(B, L, H, W, rl, seed=200 + s, pseed=s, tx=20, ty=10)
        ref = O.hetero_fusion(*inp, P, cfg)
        res[f"seed{s}_rel_l2"] = rel_l2(y, ref)
        res[f"seed{s}_max_rel"] = max_rel(y, ref)
        worst = max(worst, res[f"seed{s}_rel_l2"])
    res["worst_rel_l2"] = worst
    assert worst < 1e-3, res
    return res


def check_index_probe():
    """Bit-exact check of the CUDA window / grid index arithmetic (group_token + the relative-position offsets) against
    the reference's einops tables (tests/golden/index.npz) at 48x176, 128x128 and 96x352, through hmvit_group_attn itself:
    one agent, identity pose, Q = K = 0 and a bias table that is 0 at ONE relative offset (dr, dc) and -20000 elsewhere,
    so every query whose partner slot (qr - dr, qc - dc) exists attends to exactly that key with probability 1.0 and
    returns the partner's V row, which encodes its flat token index in 0 / 1 channels -- exact in bf16."""
    p = pkg()
    ops = p.ops
    g = np.load(os.path.join(GOLDEN, "index.npz"))
    res = {}
    for H, W in ((48, 176), (128, 128), (96, 352)):
        N = H * W
        tok = torch.arange(N)
        code = torch.zeros(N, 256)
        for c in range(32):                                   # 17 index bits + 15 bits of a hash, repeated for every head
            bits = ((tok >> c) & 1) if c < 17 else (((tok * 2654435761) >> (c - 9)) & 1)
            for h in range(8):
                code[:, h * 32 + c] = bits.float()
        v = torch.stack([code, code]).to(torch.bfloat16).to(DEV)                       # both ego-type planes
        q = torch.zeros(N, 256, dtype=torch.bfloat16, device=DEV)
        k = torch.zeros(2, N, 256, dtype=torch.bfloat16, device=DEV)
        zb = torch.zeros(2, 2, 256, device=DEV)
        geo = dict(B=1, L=1, H=H, W=W, cell=1.6, mode=torch.zeros(1, dtype=torch.int32, device=DEV),
                   record_len=torch.ones(1, dtype=torch.int32, device=DEV), cav_mask=torch.ones(1, dtype=torch.int32, device=DEV),
                   T=torch.eye(4, device=DEV).reshape(1, 1, 1, 4, 4).contiguous())
        for kind, kname in ((0, "window"), (1, "grid")):
            table_idx = torch.from_numpy(g[f"{kname}_{H}x{W}"]).long()                   # (G, 64) flat token of (group, slot)
            for dr, dc in ((0, 1), (1, 0), (-2, 3)):
                bias = torch.full((225, 8), -20000.0)
                bias[(dr + 7) * 15 + (dc + 7)] = 0.0
                # partner slot of query slot (qr, qc): (qr - dr, qc - dc)
                qs = torch.arange(64)
                pr_, pc_ = qs // 8 - dr, qs % 8 - dc
                ok = (pr_ >= 0) & (pr_ < 8) & (pc_ >= 0) & (pc_ < 8)
                expect_tok = table_idx[:, (pr_ * 8 + pc_).clamp(0, 63)]               # (G, 64)
                q_tok = table_idx[:, ok]                                             # queries that have a partner
                e_tok = expect_tok[:, ok]
                for impl in ("fused", "split", "single"):
                    if impl != "fused" and (H, W) != (48, 176):
                        continue
                    out = torch.zeros(N, 256, dtype=torch.bfloat16, device=DEV)
                    ops.group_attn(kind=kind, q=q, k=k, v=v, bk=zb, bv=zb, bias_table=bias.to(DEV), out=out, impl=impl, **geo)
                    got = out.float().cpu()
                    bad = int((got[q_tok.reshape(-1)] != code[e_tok.reshape(-1)]).any(-1).sum())
                    res[f"{H}x{W}_{kname}_{dr}_{dc}_{impl}"] = bad
                    assert bad == 0, res
    torch.cuda.synchronize()
    return res


def check_train_grads_48x176():
    """Gradient parity at the BASELINE map size (one scene, 3 mixed agents, 256 x 48 x 176) against torch.autograd through
    the fp32 CPU oracle; same stated tolerance as check_train_grads_small (1.5e-2 per parameter tensor, 5e-3 for dL/dx, forward 1e-3)."""
    r = _train_case(B=1, L=3, H=48, W=176, record_len=[3], seed=47, mode=[[1, 0, 1]])
    assert r["fwd_rel_l2"] < 1e-3 and r["dx_rel_l2"] < 5e-3 and r["param_worst_rel_l2"] < 1.5e-2 and r["params_checked"] > 30, r
    return r


CHECKS.update({"fusion_r2_goldens": check_fusion_r2_goldens, "seed_sweep": check_seed_sweep, "index_probe": check_index_probe,
               "train_grads_48x176": check_train_grads_48x176})



def check_dropout_mask():
    """hmvit_dropout: the exported CUDA mask equals the numpy Philox4x32-10 restatement bit for bit (active agents), the
    keep rate is within 4 sigma of 1 - p, out = resid + mask * a, padded slots untouched."""
    ops = pkg().ops
    B, L, N, p, seed = 2, 3, 1024, 0.1, 987654321
    rl = torch.tensor([3, 2], dtype=torch.int32, device=DEV)
    res = {}
    for stream in (0, 5):
        m = torch.full((B * L, 256, N), -1.0, device=DEV)
        ops.dropout(None, m, B=B, L=L, N=N, record_len=rl, seed=seed, stream_id=stream, p=p)
        ref = EM.dropout_mask_cm(B * L, N, seed, stream, p)
        act = [0, 1, 2, 3, 4]
        assert torch.equal(m.cpu()[act], ref[act]), stream
        assert float((m[5] + 1.0).abs().max()) == 0.0                       # padded slot of scene 1 untouched
        keep = float((m[act] != 0).float().mean())
        n = len(act) * 256 * N
        assert abs(keep - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5, keep
        res[f"keep_rate_stream{stream}"] = keep
    a = torch.randn(B * L, 256, N, device=DEV)
    r = torch.randn(B * L, 256, N, device=DEV)
    out = torch.zeros_like(a)
    ops.dropout(a, out, resid=r, B=B, L=L, N=N, record_len=rl, seed=seed, stream_id=0, p=p)
    ref = EM.dropout_mask_cm(B * L, N, seed, 0, p).to(DEV)
    assert torch.equal(out[:5], r[:5] + a[:5] * ref[:5])
    torch.cuda.synchronize()
    return res


def check_train_dropout():
    """Train mode with the shipped yaml's drop_out = 0.1 (hetero_fusion.py:66, base_transformer.py:186-190): forward and
    gradients of the CUDA path against torch.autograd through the fp32 oracle with the SAME masks replayed (exported through
    the numpy restatement of the Philox stream, which check_dropout_mask pins to the kernel bit for bit)."""
    cfg = O.default_config()
    assert cfg["hetero_fusion_block"]["drop_out"] == 0.1
    P = O.synth_state_dict(cfg, 0)
    net = pkg().HeteroFusion(cfg)
    net.load_state_dict(P, strict=True)
    net = net.to(DEV).train()
    seed = 20261017
    net.hetero_fusion_block.dropout_seed = seed
    B, L, H, W, rl = 2, 3, 16, 24, [3, 2]
    x, T, md, rlt, mask = _scene(B, L, H, W, rl, 51, tx=10, ty=5)
    g_out = torch.randn(B, 256, H, W, generator=torch.Generator().manual_seed(52))
    tr = pkg().training
    masks = []
    for s in range(2 * cfg["num_iters"]):
        d = {}
        for site, name in enumerate(("att", "hid", "ffn")):
            m = EM.dropout_mask_cm(B * L, H * W, seed, tr._drop_stream(s, site), 0.1)
            d[name] = m.view(B, L, 256, H, W).permute(0, 1, 3, 4, 2).contiguous()
        masks.append(d)
    Pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    xr = x.clone().requires_grad_(True)
    y_ref = O.hetero_fusion(xr, T, md, rlt, mask, Pg, cfg, drop_masks=masks)
    names = [k for k, v in Pg.items() if v.is_floating_point()]
    gs = torch.autograd.grad((y_ref * g_out).sum(), [xr] + [Pg[k] for k in names], allow_unused=True)
    g_ref = dict(zip(["x"] + names, gs))
    xd = x.to(DEV).requires_grad_(True)
    y = net(xd, T.to(DEV), md.to(DEV), rlt.to(DEV), mask.to(DEV))
    (y * g_out.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    res = {"fwd_rel_l2": rel_l2(y.detach().cpu(), y_ref.detach()), "dx_rel_l2": rel_l2(xd.grad.cpu(), g_ref["x"])}
    worst = 0.0
    gnorm = max(float(v.norm()) for k, v in g_ref.items() if v is not None and k != "x")
    for name, prm in net.named_parameters():
        ref = g_ref.get(name)
        if ref is None or "aggregate_fc" in name or float(ref.norm()) < 1e-4 * gnorm:
            continue
        worst = max(worst, rel_l2(prm.grad.cpu(), ref))
    res["param_worst_rel_l2"] = worst
    y_eval = O.hetero_fusion(x, T, md, rlt, mask, P, cfg)
    res["train_vs_eval_rel_l2"] = rel_l2(y_ref.detach(), y_eval)
    assert res["fwd_rel_l2"] < 1e-3 and res["dx_rel_l2"] < 3e-2 and worst < 3e-2 and res["train_vs_eval_rel_l2"] > 0.05, res
    return res


def check_train_half_autocast():
    """`--half` of the reference (train_camera.py:143-144, 178, 195-197): the module inside torch.autocast(fp16) with a
    GradScaler -- the scaled loss goes through the hand-written backward (bf16 / tf32 operands keep fp32's exponent range),
    the unscaled gradients equal the plain fp32-loss gradients."""
    cfg = O.default_config()
    cfg["hetero_fusion_block"]["drop_out"] = 0.0
    P = O.synth_state_dict(cfg, 0)
    net = pkg().HeteroFusion(cfg)
    net.load_state_dict(P, strict=True)
    net = net.to(DEV).train()
    x, T, md, rl, mask = _scene(1, 3, 16, 24, [3], 61, tx=10, ty=5)
    args = [t.to(DEV) for t in (T, md, rl, mask)]
    g_out = torch.randn(1, 256, 16, 24, generator=torch.Generator().manual_seed(62)).to(DEV)
    xd = x.to(DEV)
    y = net(xd, *args)
    (y * g_out).sum().backward()
    ref = {n: p_.grad.clone() for n, p_ in net.named_parameters() if p_.grad is not None}
    net.zero_grad(set_to_none=True)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
    opt = torch.optim.SGD([p_ for p_ in net.parameters()], lr=0.0)
    with torch.autocast("cuda", dtype=torch.float16):
        y2 = net(xd.half(), *args)                               # encoders hand fp16 features over under autocast
        loss = (y2.float() * g_out).sum()
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    worst = 0.0
    for n, p_ in net.named_parameters():
        if n in ref and float(ref[n].norm()) > 0:
            worst = max(worst, rel_l2(p_.grad.cpu(), ref[n].cpu()))
    assert torch.isfinite(y2).all() and worst < 2e-2, worst          # fp16 input rounding only
    return {"half_vs_fp32_grad_worst_rel_l2": worst, "out_dtype": str(y2.dtype)}


CHECKS.update({"dropout_mask": check_dropout_mask, "train_dropout": check_train_dropout,
               "train_half_autocast": check_train_half_autocast})


def check_fp16_feature_boundary():
    """The stated fp16-features boundary VARIANT (bench.py e2e.fp16_feature_boundary): the per-agent BEV features are handed
    over as fp16 (11-bit significand) and widened on the device.  Own parity statement: rel-L2 <= 1e-3 against the fp32
    oracle on the fp32 features (config-2 scene shape)."""
    cfg, P, net = _mk_module(0)
    x, T, md, rl, mask = _scene(1, 5, 48, 176, [5], 1236)
    with torch.no_grad():
        y16 = net(x.half().to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
        y32 = net(x.to(DEV), T.to(DEV), md.to(DEV), rl.to(DEV), mask.to(DEV)).cpu()
    ref = O.hetero_fusion(x, T, md, rl, mask, P, cfg)
    res = {"fp16_boundary_rel_l2": rel_l2(y16, ref), "fp32_boundary_rel_l2": rel_l2(y32, ref), "fp16_vs_fp32_rel_l2": rel_l2(y16, y32)}
    assert res["fp16_boundary_rel_l2"] < 1e-3, res
    return res


CHECKS.update({"fp16_feature_boundary": check_fp16_feature_boundary})


# ------------------------------------------------------------------------------------------------
# BEV encoders and BASELINE config 3
# ------------------------------------------------------------------------------------------------
def _config3_parts():
    import enc_synth as S
    sys.path.insert(0, GOLDEN)
    import make_golden_encoders as G
    enc = pkg().encoders
    args = enc.config3_args(bev_h=G.BEV_H, bev_w=G.BEV_W, image=G.IMAGE)
    la = args['lidar']
    nx, ny, _ = la['point_pillar_scatter']['grid_size']
    vox = S.synth_voxels(2, nx, ny, 400, la['lidar_range'], la['voxel_size'], seed=1)
    cams = S.synth_cameras(3, 2, G.IMAGE, seed=2)
    return S, G, enc, args, vox, cams, np.load(os.path.join(GOLDEN, "encoders.npz"))


def check_encoders_golden():
    """PointPillar and the CVT camera branch (library-backed torch modules, hm-vit_b200/encoders.py) ON THE GPU against the
    outputs of the unmodified reference modules.  cuDNN / cuBLAS TF32 is switched off for the comparison (fp32 like the
    reference on CPU); stated tolerance rel-L2 <= 1e-3 per tensor."""
    S, G, enc, args, vox, cams, gold = _config3_parts()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        pp = enc.PointPillar(args['lidar']).eval()
        pp.load_state_dict(S.synth_module_state_dict(pp, 1), strict=True)
        pp = pp.to(DEV).set_return_features()
        cam = enc.CvtCameraEncoder(args['camera']).eval()
        cam.encoder.load_state_dict(S.synth_module_state_dict(cam.encoder, 2), strict=True)
        cam.cvm.load_state_dict(S.synth_module_state_dict(cam.cvm, 3), strict=True)
        cam = cam.to(DEV)
        with torch.no_grad():
            lf = pp({'processed_lidar': {k: v.to(DEV) for k, v in vox.items()}, 'batch_size': 2})
            cams_d = {k: v.to(DEV) for k, v in cams.items()}
            cf_fused = cam(cams_d)                      # default on CUDA: per-camera fused attention + log-sum-exp merge
            for cv in cam.cvm.cross_views:
                cv.cross_attend.fused = False
            cf = cam(cams_d)                            # materialised fp32 logits, the reference's form
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    res = {"point_pillar_rel_l2": rel_l2(lf.cpu(), torch.from_numpy(gold['pp_features'])),
           "cvt_rel_l2": rel_l2(cf.cpu(), torch.from_numpy(gold['cvt_features'])),
           "cvt_fused_attention_rel_l2": rel_l2(cf_fused.cpu(), torch.from_numpy(gold['cvt_features']))}
    assert res["point_pillar_rel_l2"] < 1e-3 and res["cvt_rel_l2"] < 1e-3 and res["cvt_fused_attention_rel_l2"] < 1e-3, res
    return res


def check_config3_golden():
    """BASELINE config 3 end to end on the GPU: raw voxels + camera images -> encoders -> combine / regroup -> HeteroFusion
    (hmvit_fusion_forward) -> HeteroDecoder (hmvit_decoder_forward) -> psm / rm through `BevformerPointPillarHetero.forward(batch)`,
    against the same pipeline assembled from the UNMODIFIED reference modules (tests/golden/encoders.npz).  Stated tolerance
    for this composed case: rel-L2 <= 2e-3 per tensor (see check_model_glue)."""
    S, G, enc, args, vox, cams, gold = _config3_parts()
    mode = torch.tensor(G.C3_MODE)
    rl = torch.tensor(G.C3_RECORD_LEN)
    cfgf = O.default_config()
    args = dict(args, hetero_fusion=cfgf, spatial_transform=cfgf["spatial_transform"], max_cav=3)
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = enc.build_config3_model(args).eval()
        net.lidar_encoder.load_state_dict(S.synth_module_state_dict(net.lidar_encoder, 1), strict=True)
        net.camera_encoder.encoder.load_state_dict(S.synth_module_state_dict(net.camera_encoder.encoder, 2), strict=True)
        net.camera_encoder.cvm.load_state_dict(S.synth_module_state_dict(net.camera_encoder.cvm, 3), strict=True)
        net.fusion_net.load_state_dict(O.synth_state_dict(cfgf, 0), strict=True)
        net.decoder.load_state_dict(O.synth_decoder_state_dict(1), strict=True)
        net = net.to(DEV)
        batch = S.config3_batch(vox, cams, G.C3_ORDER, mode, rl, torch.from_numpy(gold['c3_T']))

        def to_dev(v):
            return {k: to_dev(x) for k, x in v.items()} if isinstance(v, dict) else v.to(DEV)
        with torch.no_grad():
            out = net(to_dev(batch))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    res = {"psm_rel_l2": rel_l2(out["psm"].cpu(), torch.from_numpy(gold['c3_psm'])),
           "rm_rel_l2": rel_l2(out["rm"].cpu(), torch.from_numpy(gold['c3_rm']))}
    assert res["psm_rel_l2"] < 2e-3 and res["rm_rel_l2"] < 2e-3, res
    return res


def check_pillar_scatter():
    """hmvit_pillar_scatter (PillarVFE + PointPillarScatter in one kernel, csrc/pillar.cuh) against the torch modules of the
    same PointPillar (themselves pinned to the reference, test_encoders_cpu.py): the dense canvas, for pillars with 1..32
    points incl. FULL pillars (no padded slot: the maximum may be negative before the ReLU), an agent without any pillar, and
    through the whole PointPillar vs the reference golden.  Bar: max abs difference <= 1e-5 on the canvas (fp32, different
    accumulation order of the 10 products)."""
    S, G, enc, args, vox, cams, gold = _config3_parts()
    la = args['lidar']
    pp = enc.PointPillar(la).eval()
    pp.load_state_dict(S.synth_module_state_dict(pp, 1), strict=True)
    pp = pp.to(DEV).set_return_features()
    nx, ny, _ = la['point_pillar_scatter']['grid_size']
    v = S.synth_voxels(3, nx, ny, 500, la['lidar_range'], la['voxel_size'], seed=9)
    keep = v['voxel_coords'][:, 0] != 1                                  # agent 1 has no pillar at all
    v = {k: t[keep] for k, t in v.items()}
    v['voxel_num_points'][::7] = 32                                      # full pillars
    full = v['voxel_num_points'] == 32
    g = torch.Generator().manual_seed(3)
    v['voxel_features'][full] = v['voxel_features'][full] + (v['voxel_features'][full] == 0) * torch.rand(int(full.sum()), 32, 4, generator=g)
    vd = {k: t.to(DEV) for k, t in v.items()}
    with torch.no_grad():
        d = {'voxel_features': vd['voxel_features'], 'voxel_coords': vd['voxel_coords'], 'voxel_num_points': vd['voxel_num_points'],
             'batch_size': 3}
        fused = pp._fused_front_end(dict(d))
        assert fused is not None, "the fused front end must be taken for the shipped configuration on CUDA"
        ref = pp.scatter(pp.pillar_vfe(dict(d)))['spatial_features']
        pp.channels_last = True
        fused_cl = pp._fused_front_end(dict(d))                          # same logical tensor, channels_last memory format
        pp.channels_last = False
        assert fused_cl.is_contiguous(memory_format=torch.channels_last) and torch.equal(fused_cl, fused)
        # whole module, both paths, vs the reference golden
        gv = {k: t.to(DEV) for k, t in vox.items()}
        f1 = pp({'processed_lidar': gv, 'batch_size': 2})
        pp.fused_front_end = False
        f0 = pp({'processed_lidar': gv, 'batch_size': 2})
    gp = torch.from_numpy(gold['pp_features'])
    res = {"canvas_max_abs_diff": float((fused - ref).abs().max()), "canvas_nonzero_cells": int((ref != 0).any(1).sum()),
           "empty_agent_zero": bool((fused[1] == 0).all()), "full_pillars": int(full.sum()),
           "point_pillar_fused_rel_l2": rel_l2(f1.cpu(), gp), "point_pillar_torch_rel_l2": rel_l2(f0.cpu(), gp)}
    assert tuple(fused.shape) == tuple(ref.shape) == (3, 64, ny, nx)
    assert res["canvas_max_abs_diff"] <= 1e-5 and res["empty_agent_zero"] and res["full_pillars"] > 0, res
    assert res["point_pillar_fused_rel_l2"] < 1e-3 and res["point_pillar_torch_rel_l2"] < 1e-3, res
    return res


CHECKS.update({"encoders_golden": check_encoders_golden, "config3_golden": check_config3_golden,
               "pillar_scatter": check_pillar_scatter})
