"""TEST INFRASTRUCTURE ONLY -- torch (CPU or GPU tensor) restatement of every kernel entry point that
hmvit_b200/training.py calls, with the SAME signatures and buffer layouts as hmvit_b200/ops.py.

Two uses:
  * CPU suite: training.fusion_train(emul_ops, ...) runs the product's orchestration + backward algebra in
    exact fp32 arithmetic (ROWS_DTYPE = float32, no bf16 / tf32 rounding) and is compared with autograd of the
    oracle (tests/test_training_cpu.py);
  * GPU suite: every CUDA kernel is compared op by op with the function of the same name here
    (tests/gpu_checks.py).
The attention backward is obtained by autograd of the emulated attention forward, i.e. it does not share the
hand-derived formulas of csrc/attn_bwd.cuh.  Never imported by the product package.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from oracle import hmvit_oracle as O

ROWS_DTYPE = torch.float32
_LOG2E = 1.4426950408889634
C = 256
(GEMM_QKV, GEMM_OUT, GEMM_FFN1, GEMM_FFN2, GEMM_HEAD1, GEMM_HEAD2, GEMM_QKV_NOLN,
 GEMM_LN_LIN_CM, GEMM_LIN_CM, GEMM_LIN_ROWS, GEMM_ROWS_LIN_CM) = range(11)


def _agents(B, L, record_len, ego_only):
    for b in range(B):
        for l in range(int(record_len[b])):
            if ego_only and l != 0:
                continue
            yield b, l, b * L + l


def _ln(v, eps=1e-5):
    return F.layer_norm(v, (v.shape[-1],), None, None, eps)


def rowgemm(variant, *, B, L, N, n_out, mode, record_len, a, w0, w1, bias, out, ego_only=False,
            ln_gamma=None, ln_beta=None, ln_eps=1e-5, resid=None, ln_stats=None):
    w = (w0.float(), w1.float())
    mode = mode.reshape(-1)
    if variant in (GEMM_QKV, GEMM_QKV_NOLN):
        for b in range(B):
            n = int(record_len[b])
            types = [int(mode[b * L + j] != 0) for j in range(n)]
            te_set = {types[0]} if ego_only else set(types)
            for l in range(n):
                ai, t = b * L + l, types[l]
                A = a[ai].t().float()
                if variant == GEMM_QKV:
                    A = _ln(A, ln_eps)
                y = A @ w[t].t() + bias[t]
                planes = ([0] if (not ego_only or l == 0) else []) + [1 + te for te in te_set] + [3 + te for te in te_set]
                for p in planes:
                    out[p, ai * N:(ai + 1) * N] = y[:, p * C:(p + 1) * C].to(out.dtype)
        return out
    head = variant in (GEMM_HEAD1, GEMM_HEAD2)
    for b, l, ai in _agents(B, L, record_len, ego_only or head):
        t = int(mode[ai] != 0)
        if variant in (GEMM_OUT, GEMM_ROWS_LIN_CM):
            A = a[ai * N:(ai + 1) * N].float()
        else:
            A = a[ai].t().float()
            if variant in (GEMM_LN_LIN_CM, GEMM_FFN1):
                A = _ln(A, ln_eps)
        y = A @ w[t].t() + bias[t]
        if resid is not None and variant in (GEMM_OUT, GEMM_ROWS_LIN_CM, GEMM_FFN2):
            y = y + resid[ai].t()
        if variant in (GEMM_HEAD1, GEMM_FFN1):
            y = F.gelu(y)
        if variant == GEMM_LIN_ROWS:
            out[ai * N:(ai + 1) * N] = y.to(out.dtype)
        elif variant == GEMM_HEAD2:
            out[b] = y.t()
        else:
            out[ai] = y.t()
    return out


def out_ffn_chain(*, B, L, N, mode, record_len, o, resid, out, wa0, wa1, ba, w1_0, w1_1, b1, w2_0, w2_1, b2,
                  ln_gamma=None, ln_beta=None, ego_only=False, ln_eps=1e-5, stats_out=None):
    wa, w1, w2 = (wa0.float(), wa1.float()), (w1_0, w1_1), (w2_0, w2_1)
    mode = mode.reshape(-1)
    for b, l, ai in _agents(B, L, record_len, ego_only):
        t = int(mode[ai] != 0)
        x1 = resid[ai].t() + o[ai * N:(ai + 1) * N].float() @ wa[t].t() + ba[t]
        h = F.gelu(_ln(x1, ln_eps) @ w1[t].t() + b1[t])
        out[ai] = (x1 + h @ w2[t].t() + b2[t]).t()
        if stats_out is not None:                                    # (mean, rstd) of every output row, for the next stage's LayerNorm
            v = out[ai]
            st = stats_out.view(B * L, N, 2)
            st[ai, :, 0] = v.mean(0)
            st[ai, :, 1] = torch.rsqrt(v.var(0, unbiased=False) + ln_eps)
    return out


# ---- attention ---------------------------------------------------------------------------------------
def _attn_scene_ego(b, i, n, L, H, W, kind, mode, cav_mask, T, cell, q, k, v, bk, bv, bias_table):
    """Differentiable emulated attention of ego i of scene b: returns (out (N, 256), lse2 (N, 8)); q (R, 256),
    k / v (2, R, 256) may require grad."""
    N = H * W
    h, d = 8, 32
    mode = mode.reshape(-1)
    table = O.group_token_table(H, W, 8, "window" if kind == 0 else "grid").to(q.device)      # (G, S)
    G, S = table.shape
    rel = O.relative_position_index(8).to(q.device)
    te = int(mode[b * L + i] != 0)
    src = torch.arange(n)
    Tb = T.view(-1, L, L, 4, 4)[b, :n, i].cpu()
    sx, sy = O.source_coords(Tb, H, W, cell, 1.0)
    rx, ry = torch.round(sx), torch.round(sy)
    vis = ((rx >= 0) & (rx <= W - 1) & (ry >= 0) & (ry <= H - 1)) & (cav_mask.view(-1, L)[b, :n].cpu().view(n, 1, 1) != 0)
    vis = vis.reshape(n, N).to(q.device)
    sx, sy = sx.to(q.device), sy.to(q.device)
    rows = torch.stack([torch.arange((b * L + int(j)) * N, (b * L + int(j) + 1) * N) for j in src]).to(q.device)
    Kp = k[te][rows].float().view(n, H, W, C)
    Vp = v[te][rows].float().view(n, H, W, C)
    tj = [int(mode[b * L + int(j)] != 0) for j in src]
    bks = torch.stack([bk[te, t] for t in tj]).view(n, 1, C)
    bvs = torch.stack([bv[te, t] for t in tj]).view(n, 1, C)
    Kg = (O.warp_bilinear_nhwc(Kp, sx, sy).reshape(n, N, C) + bks)[:, table].view(n, G, S, h, d)
    Vg = (O.warp_bilinear_nhwc(Vp, sx, sy).reshape(n, N, C) + bvs)[:, table].view(n, G, S, h, d)
    qg = q[(b * L + i) * N:(b * L + i + 1) * N].float()[table].view(G, S, h, d)
    bias2 = bias_table[rel].permute(2, 0, 1) * _LOG2E                                            # (h, S, S)
    logits = torch.einsum("gshd,jgkhd->ghsjk", qg, Kg) + bias2[None, :, :, None, :]
    km = vis[:, table].permute(1, 0, 2).reshape(G, 1, 1, n, S)
    logits = logits.masked_fill(~km, float("-inf")).reshape(G, h, S, n * S)
    lse2 = torch.logsumexp(logits * math.log(2.0), dim=-1) * _LOG2E                               # (G, h, S)
    p = torch.exp2(logits - lse2[..., None]).view(G, h, S, n, S)
    o = torch.einsum("ghsjk,jgkhd->gshd", p, Vg).reshape(G * S, C)
    out = torch.zeros(N, C, dtype=o.dtype, device=o.device).index_copy(0, table.reshape(-1), o)
    lse = torch.zeros(N, h, dtype=o.dtype, device=o.device).index_copy(0, table.reshape(-1), lse2.permute(0, 2, 1).reshape(G * S, h))
    return out, lse


def group_attn(*, B, L, H, W, kind, mode, record_len, cav_mask, T, cell, q, k, v, bk, bv, bias_table, out,
               ego_only=False, key_mask=None, lse=None):
    assert key_mask is None
    N = H * W
    with torch.no_grad():
        for b, i, ai in _agents(B, L, record_len, ego_only):
            o, l2 = _attn_scene_ego(b, i, int(record_len[b]), L, H, W, kind, mode, cav_mask, T, cell, q, k, v, bk, bv, bias_table)
            out[ai * N:(ai + 1) * N] = o.to(out.dtype)
            if lse is not None:
                lse[ai * N:(ai + 1) * N] = l2
    return out


def group_attn_bwd(*, B, L, H, W, kind, mode, record_len, cav_mask, T, cell, q, k, v, bk, bv, bias_table, o, d_o, lse,
                   dq, dk, dv, dbk, dbv, dbias_table, ego_only=False):
    """autograd of the emulated forward; gradients ACCUMULATE into dq / dk / dv / dbk / dbv / dbias_table."""
    N = H * W
    with torch.enable_grad():
        leaves = [t.detach().float().clone().requires_grad_(True) for t in (q, k, v, bk, bv, bias_table)]
        tot = None
        for b, i, ai in _agents(B, L, record_len, ego_only):
            oo, _ = _attn_scene_ego(b, i, int(record_len[b]), L, H, W, kind, mode, cav_mask, T, cell, *leaves)
            term = (oo * d_o[ai * N:(ai + 1) * N].float()).sum()
            tot = term if tot is None else tot + term
        gs = torch.autograd.grad(tot, leaves, allow_unused=True)
    for dst, g in zip((dq, dk, dv, dbk, dbv, dbias_table), gs):
        if g is not None:
            dst += g.to(dst.dtype)


# ---- small backward kernels --------------------------------------------------------------------------
def bwd_row_stats(x, stats, *, B, L, N, record_len, ego_only=False, eps=1e-5):
    st = stats.view(B * L, N, 2)
    for b, l, ai in _agents(B, L, record_len, ego_only):
        v = x[ai]                                                    # (256, N)
        mean = v.mean(0)
        var = v.var(0, unbiased=False)
        st[ai, :, 0] = mean
        st[ai, :, 1] = torch.rsqrt(var + eps)
    return stats


def bwd_layernorm(dz, x, stats, dres, dx, *, B, L, N, record_len, ego_only=False):
    st = stats.view(B * L, N, 2)
    for b, l, ai in _agents(B, L, record_len, ego_only):
        mean, rstd = st[ai, :, 0], st[ai, :, 1]
        z = (x[ai] - mean) * rstd
        g = dz[ai]
        dx[ai] = dres[ai] + rstd * (g - g.mean(0) - z * (g * z).mean(0))
    return dx


def bwd_gelu(hp, dh):
    v = hp.clone()
    cdf = 0.5 * (1.0 + torch.erf(v * 0.7071067811865476))
    pdf = 0.3989422804014327 * torch.exp(-0.5 * v * v)
    hp.copy_(v * cdf)
    dh.mul_(cdf + v * pdf)


def bwd_cast_bf16(src, dst):
    dst.copy_(src.to(dst.dtype))
    return dst


def bwd_colsum(y, db, *, B, L, N, mode, record_len, ego_only=False):
    rows = y.dim() == 2
    mode = mode.reshape(-1)
    for b, l, ai in _agents(B, L, record_len, ego_only):
        t = int(mode[ai] != 0)
        s = y[ai * N:(ai + 1) * N].float().sum(0) if rows else y[ai].sum(1)
        db[t, :C] += s
    return db


def _operand(t, ai, N):
    """(N tokens, 256) view of agent ai from a cm (B*L, 256, N) or rows (R, 256) tensor."""
    return t[ai * N:(ai + 1) * N].float() if t.dim() == 2 else t[ai].t()


def bwd_cast_colsum(src, dst, db, *, B, L, N, mode):
    dst.copy_(src.to(dst.dtype))
    m = mode.reshape(-1)
    for ai in range(B * L):
        db[int(m[ai] != 0), :5 * C] += src[:, ai * N:(ai + 1) * N].float().sum(1).reshape(-1)
    return dst


def bwd_dgrad_cat(dcat, w0, w1, out, *, B, L, N, mode, record_len):
    w = (w0.float(), w1.float())
    mode = mode.reshape(-1)
    for b, l, ai in _agents(B, L, record_len, False):
        t = int(mode[ai] != 0)
        y = sum(dcat[p, ai * N:(ai + 1) * N].float() @ w[t][p * C:(p + 1) * C].t() for p in range(5))
        out[ai] = y.t()
    return out


def bwd_wgrad(a, b, dw, *, B, L, N, mode, record_len, ego_only=False, b_stats=None, row0=0):
    mode = mode.reshape(-1)
    for bb, l, ai in _agents(B, L, record_len, ego_only):
        t = int(mode[ai] != 0)
        Bm = _operand(b, ai, N)
        if b_stats is not None:
            st = b_stats.view(B * L, N, 2)[ai]
            Bm = (Bm - st[:, :1]) * st[:, 1:]
        dw[t, row0:row0 + C] += _operand(a, ai, N).t() @ Bm
    return dw


# ---- dropout: the same counter-based mask as csrc/dropout.cuh (Philox4x32-10), in numpy --------------------
def _philox4x32_10(c0, c1, c2, c3, k0, k1):
    import numpy as np
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [c0.astype(np.uint64), c1.astype(np.uint64), c2.astype(np.uint64), c3.astype(np.uint64)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    mask32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask32, p1 >> np.uint64(32), p1 & mask32
        c = [(hi1 ^ c[1] ^ k0) & mask32, lo1, (hi0 ^ c[3] ^ k1) & mask32, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask32
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask32
    return c


def dropout_mask_cm(agents, N, seed, stream_id, p):
    """Scaled keep mask (agents, 256, N) fp32 = what hmvit_dropout(a=NULL) writes for active agents."""
    import numpy as np
    assert N % 4 == 0
    nblk = agents * C * N // 4
    idx = np.arange(nblk, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = _philox4x32_10(idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), np.full(nblk, int(stream_id) & 0xFFFFFFFF, dtype=np.uint64),
                       np.zeros(nblk, dtype=np.uint64), seed & 0xFFFFFFFF, seed >> 32)
    r = np.stack(r, axis=1).reshape(-1)                        # 4 consecutive elements per block
    th = min(int(float(np.float32(p)) * 4294967296.0), 4294967295)
    keep = (r >= np.uint64(th)).astype(np.float32) * np.float32(1.0 / (1.0 - float(np.float32(p))))
    return torch.from_numpy(keep.reshape(agents, C, N))


def dropout(a, out, *, B, L, N, record_len, seed, stream_id, p, resid=None, ego_only=False):
    m = dropout_mask_cm(B * L, N, seed, stream_id, p).to(out.device)
    for b, l, ai in _agents(B, L, record_len, ego_only):
        v = m[ai] if a is None else a[ai] * m[ai]
        out[ai] = v + (resid[ai] if resid is not None else 0.0)
    return out
