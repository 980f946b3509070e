"""Training path of the HM-ViT fusion module: forward that saves activations + hand-written backward.

The reference trains the fusion module with autograd (opencood/tools/train_camera.py:172-193 around
`model(batch_data['ego'])`, i.e. HeteroFusion.forward at bevformer_point_pillar_hetero.py:39-49).  Here the
forward runs the same sm_100a kernels as inference, stage by stage, keeping per stage the stage input
(fp32), the projected Q / K' / V' rows, the attention output and the softmax log-sum-exp; the backward is the
adjoint of the RESTRUCTURED forward:

  * input-gradient GEMMs  = the tcgen05 row-GEMM with transposed weights (ops.rowgemm variants 7-10); the five planes of the
                            fused Q | K' | V' projection as ONE K = 1280 GEMM (ops.bwd_dgrad_cat);
  * weight gradients      = ops.bwd_wgrad (typed, tcgen05 with the accumulator in tensor memory), bias gradients =
                            ops.bwd_colsum / ops.bwd_cast_colsum (the projection's, in the pass that casts its gradient planes);
  * LayerNorm / GELU      = ops.bwd_layernorm / ops.bwd_gelu;
  * attention             = ops.group_attn_bwd (re-gathers K'/V', scatters dK'/dV' through the bilinear taps).

The kernels produce gradients of the FOLDED weights (edge-type weights, softmax scale and LayerNorm affines
folded into the projections); `fold_stage` / `fold_head` below are the differentiable folding functions and
torch.autograd pulls the folded gradients back to the module parameters (a few 256 x 256 products).

`ops` is passed in explicitly: the product path always passes hmvit_b200.ops (CUDA, no fallback); the CPU test
suite passes tests/emul_ops.py, a torch restatement of every kernel, to check this orchestration and the
backward algebra against autograd of the oracle without a GPU.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib

_LOG2E = 1.4426950408889634
C_DIM = 256


# ----------------------------------------------------------------------------------------------
# differentiable folding (must stay in lock step with HeteroAttention.packed / HeteroFusionBlock._stage_pack)
# ----------------------------------------------------------------------------------------------
def fold_stage(block, kind: str) -> Dict[str, torch.Tensor]:
    """fp32 folded weights of one stage (window | grid); differentiable w.r.t. the module parameters."""
    att = getattr(block, f"{kind}_attention")
    norm = getattr(block, f"{kind}_norm")
    ffd = getattr(block, f"{kind}_ffd")
    Cd, h, d = att._dim, att.heads, att._dim_head
    ratt, rmsg = att.relation_att.float(), att.relation_msg.float()
    wcat, bcat, wa, ba, w1, b1, w2, b2 = [], [], [], [], [], [], [], []
    bk = [[None, None], [None, None]]
    bv = [[None, None], [None, None]]
    for t in range(2):
        g1, be1 = norm.net[t].weight.float(), norm.net[t].bias.float()
        wq = att.q_linears[t].weight.float() * (att.scale * _LOG2E)
        bq = att.q_linears[t].bias.float() * (att.scale * _LOG2E)
        wk = att.k_linears[t].weight.float().view(h, d, Cd)
        bkt = att.k_linears[t].bias.float().view(h, d)
        wv = att.v_linears[t].weight.float().view(h, d, Cd)
        bvt = att.v_linears[t].bias.float().view(h, d)
        parts = [wq]
        for te in range(2):
            e = te * 2 + t
            parts.append(torch.einsum("hpq,hqc->hpc", ratt[e], wk).reshape(Cd, Cd))
            bk[te][t] = torch.einsum("hpq,hq->hp", ratt[e], bkt).reshape(Cd)
        for te in range(2):
            e = te * 2 + t
            parts.append(torch.einsum("hpq,hpc->hqc", rmsg[e], wv).reshape(Cd, Cd))
            bv[te][t] = torch.einsum("hpq,hp->hq", rmsg[e], bvt).reshape(Cd)
        wc = torch.cat(parts, 0)
        bc = torch.cat([bq, bq.new_zeros(4 * Cd)])
        bcat.append(bc + wc @ be1)
        wcat.append(wc * g1[None, :])
        wa.append(att.a_linears[t][0].weight.float())
        ba.append(att.a_linears[t][0].bias.float())
        g2, be2 = ffd.norm.net[t].weight.float(), ffd.norm.net[t].bias.float()
        w1t = ffd.fn.net[t][0].weight.float()
        w1.append(w1t * g2[None, :])
        b1.append(ffd.fn.net[t][0].bias.float() + w1t @ be2)
        w2.append(ffd.fn.net[t][3].weight.float())
        b2.append(ffd.fn.net[t][3].bias.float())
    return {
        "wcat": torch.stack(wcat), "bcat": torch.stack(bcat),
        "bk": torch.stack([torch.stack(r) for r in bk]), "bv": torch.stack([torch.stack(r) for r in bv]),
        "wa": torch.stack(wa), "ba": torch.stack(ba), "w1": torch.stack(w1), "b1": torch.stack(b1),
        "w2": torch.stack(w2), "b2": torch.stack(b2),
        "bias_table": att.relative_position_bias_table.weight.float(),
    }


def fold_head(fusion) -> Dict[str, torch.Tensor]:
    net = fusion.mlp_head.net
    return {"w1": torch.stack([net[t][0].weight.float() for t in range(2)]),
            "b1": torch.stack([net[t][0].bias.float() for t in range(2)]),
            "w2": torch.stack([net[t][3].weight.float() for t in range(2)]),
            "b2": torch.stack([net[t][3].bias.float() for t in range(2)])}


FOLD_KEYS = ("wcat", "bcat", "bk", "bv", "wa", "ba", "w1", "b1", "w2", "b2", "bias_table")
HEAD_KEYS = ("w1", "b1", "w2", "b2")


def _tf32(w: torch.Tensor, exact: bool) -> torch.Tensor:
    w = w.detach().float().contiguous()
    if exact:
        return w
    return ((w.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def _lowp(w: torch.Tensor, rows_dtype) -> torch.Tensor:
    return w.detach().to(rows_dtype).contiguous()


def kernel_pack_stage(F: Dict[str, torch.Tensor], rows_dtype) -> Dict[str, torch.Tensor]:
    """Kernel-side copies of a folded stage: forward operands (same as inference) + transposed copies for the
    input-gradient GEMMs.  rows_dtype is bf16 on the GPU; the CPU emulation passes fp32 (exact arithmetic)."""
    exact = rows_dtype == torch.float32
    Cd = C_DIM
    pk = {}
    for t in range(2):
        pk[f"wqkv{t}"] = _lowp(F["wcat"][t], rows_dtype)
        pk[f"wa{t}"] = _lowp(F["wa"][t], rows_dtype)
        pk[f"w1_{t}"] = _tf32(F["w1"][t], exact)
        pk[f"w2_{t}"] = _tf32(F["w2"][t], exact)
        # the fused chain kernel takes fp16 feed-forward weights (same significand width as tf32); exact on the CPU emulation
        pk[f"w1h_{t}"] = F["w1"][t].detach().contiguous() if exact else F["w1"][t].detach().half().contiguous()
        pk[f"w2h_{t}"] = F["w2"][t].detach().contiguous() if exact else F["w2"][t].detach().half().contiguous()
        pk[f"waT{t}"] = _tf32(F["wa"][t].t(), exact)
        pk[f"w1T{t}"] = _tf32(F["w1"][t].t(), exact)
        pk[f"w2T{t}"] = _tf32(F["w2"][t].t(), exact)
        # input gradient of the fused projection as one K = 1280 GEMM: rows p*256.. = transposed plane p of W_cat
        pk[f"wcatT_{t}"] = _lowp(torch.cat([F["wcat"][t][p * Cd:(p + 1) * Cd].t() for p in range(5)], 0), rows_dtype)
    for k in ("bcat", "bk", "bv", "ba", "b1", "b2", "bias_table"):
        pk[k] = F[k].detach().float().contiguous()
    return pk


def kernel_pack_head(F: Dict[str, torch.Tensor], rows_dtype) -> Dict[str, torch.Tensor]:
    exact = rows_dtype == torch.float32
    pk = {}
    for t in range(2):
        pk[f"w1_{t}"] = _tf32(F["w1"][t], exact)
        pk[f"w2_{t}"] = _tf32(F["w2"][t], exact)
        pk[f"w1T{t}"] = _tf32(F["w1"][t].t(), exact)
        pk[f"w2T{t}"] = _tf32(F["w2"][t].t(), exact)
    pk["b1"] = F["b1"].detach().float().contiguous()
    pk["b2"] = F["b2"].detach().float().contiguous()
    return pk


# ----------------------------------------------------------------------------------------------
# forward with saved activations
# ----------------------------------------------------------------------------------------------
class _Saved:
    pass


def _drop_stream(stage: int, site: int) -> int:
    """Philox stream of a Dropout site: site 0 = behind a_linears (hetero_fusion.py:66), 1 = behind the FFN's GELU,
    2 = behind the FFN's second Linear (base_transformer.py:186-190)."""
    return stage * 4 + site


def forward_train(ops, geo, x, packs, head_pack, num_iters, skip_dead, drop_p=0.0, seed=0):
    """Stage-by-stage forward through the per-kernel entry points.  Returns (out | None, last x, saved).
    drop_p > 0 (train mode of the shipped yaml): the three Dropout sites of a stage sit between the GEMMs, so the stage
    runs as row-GEMMs + the Philox dropout kernel instead of the fused chain kernel; masks are regenerated, not saved."""
    B, L, H, W = geo["B"], geo["L"], geo["H"], geo["W"]
    N = H * W
    R = B * L * N
    dev = x.device
    rows_dtype = getattr(ops, "ROWS_DTYPE", torch.bfloat16)
    mode, rl, cav, T, cell = geo["mode"], geo["record_len"], geo["cav_mask"], geo["T"], geo["cell"]
    common = dict(B=B, L=L, N=N, mode=mode, record_len=rl)
    sv = _Saved()
    sv.xs, sv.qkv, sv.att, sv.lse, sv.dead = [x], [], [], [], []
    sv.stats = [None]                        # LayerNorm statistics of sv.xs[s] when the producing chain launch wrote them
    head = head_pack is not None
    n_stage = 2 * num_iters
    for s in range(n_stage):
        kind = s & 1
        pk = packs[kind]
        dead = bool(head and skip_dead and s == n_stage - 1)
        xin = sv.xs[-1]
        qkv = torch.empty(5, R, C_DIM, dtype=rows_dtype, device=dev)
        ops.rowgemm(_lib.GEMM_QKV, n_out=5 * C_DIM, a=xin, w0=pk["wqkv0"], w1=pk["wqkv1"], bias=pk["bcat"], out=qkv,
                    ego_only=dead, ln_stats=sv.stats[-1], **common)
        att = torch.empty(R, C_DIM, dtype=rows_dtype, device=dev)
        lse = torch.empty(R, 8, dtype=torch.float32, device=dev)
        ops.group_attn(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell,
                       q=qkv[0], k=qkv[1:3], v=qkv[3:5], bk=pk["bk"], bv=pk["bv"], bias_table=pk["bias_table"],
                       out=att, ego_only=dead, lse=lse)
        # the stage output is written for every active agent and read for active agents only (padded slots pass the input
        # through at the end, _FusionFn.forward): no zero fill
        xout = torch.empty_like(xin)
        st_next = None
        if drop_p > 0.0:
            dk = dict(B=B, L=L, N=N, record_len=rl, seed=seed, p=drop_p, ego_only=dead)
            tmp = torch.zeros_like(xin)
            xp = torch.zeros_like(xin)
            ops.rowgemm(_lib.GEMM_ROWS_LIN_CM, n_out=C_DIM, a=att, w0=pk["wa0"], w1=pk["wa1"], bias=pk["ba"], out=tmp,
                        ego_only=dead, **common)
            ops.dropout(tmp, xp, resid=xin, stream_id=_drop_stream(s, 0), **dk)          # x' = x + Dropout(O Wa^T + ba)
            ops.rowgemm(_lib.GEMM_FFN1, n_out=C_DIM, a=xp, w0=pk["w1_0"], w1=pk["w1_1"], bias=pk["b1"], out=tmp,
                        ego_only=dead, **common)                                        # gelu(W1' LN(x') + b1')
            ops.dropout(tmp, tmp, stream_id=_drop_stream(s, 1), **dk)
            hid = tmp
            tmp = torch.zeros_like(xin)
            ops.rowgemm(_lib.GEMM_LIN_CM, n_out=C_DIM, a=hid, w0=pk["w2_0"], w1=pk["w2_1"], bias=pk["b2"], out=tmp,
                        ego_only=dead, **common)
            ops.dropout(tmp, xout, resid=xp, stream_id=_drop_stream(s, 2), **dk)         # x'' = x' + Dropout(W2 h + b2)
            del tmp, hid, xp
        else:
            # the chain kernel hands the next stage the LayerNorm statistics of its output rows (as the inference path does):
            # no statistics pass in the next QKV launch nor in the backward of the next stage
            st_next = None if dead else torch.empty(R, 2, dtype=torch.float32, device=dev)
            ops.out_ffn_chain(o=att, resid=xin, out=xout, wa0=pk["wa0"], wa1=pk["wa1"], ba=pk["ba"],
                              w1_0=pk["w1h_0"], w1_1=pk["w1h_1"], b1=pk["b1"], w2_0=pk["w2h_0"], w2_1=pk["w2h_1"], b2=pk["b2"],
                              ego_only=dead, stats_out=st_next, **common)
        sv.stats.append(st_next)
        sv.qkv.append(qkv); sv.att.append(att); sv.lse.append(lse); sv.dead.append(dead); sv.xs.append(xout)
    out = None
    if head:
        hid = torch.empty(B * L, C_DIM, N, dtype=torch.float32, device=dev)
        out = torch.empty(B, C_DIM, N, dtype=torch.float32, device=dev)
        ops.rowgemm(_lib.GEMM_HEAD1, n_out=C_DIM, a=sv.xs[-1], w0=head_pack["w1_0"], w1=head_pack["w1_1"], bias=head_pack["b1"],
                    out=hid, **common)
        ops.rowgemm(_lib.GEMM_HEAD2, n_out=C_DIM, a=hid, w0=head_pack["w2_0"], w1=head_pack["w2_1"], bias=head_pack["b2"],
                    out=out, **common)
    return out, sv.xs[-1], sv


# ----------------------------------------------------------------------------------------------
# backward
# ----------------------------------------------------------------------------------------------
def _zero_grads(dev) -> Dict[str, torch.Tensor]:
    z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)  # noqa: E731
    return {"wcat": z(2, 5 * C_DIM, C_DIM), "bcat": z(2, 5 * C_DIM), "bk": z(2, 2, C_DIM), "bv": z(2, 2, C_DIM),
            "wa": z(2, C_DIM, C_DIM), "ba": z(2, C_DIM), "w1": z(2, C_DIM, C_DIM), "b1": z(2, C_DIM),
            "w2": z(2, C_DIM, C_DIM), "b2": z(2, C_DIM), "bias_table": z(225, 8)}


def backward(ops, geo, sv, packs, head_pack, d_out: Optional[torch.Tensor], d_xlast: Optional[torch.Tensor],
             drop_p=0.0, seed=0):
    """Returns (dx [B*L, 256, N] fp32, [grads window, grads grid], head grads | None): gradients of the folded
    weights.  d_out: gradient of the head output (B, 256, N); d_xlast: gradient of the block output (no head).
    drop_p / seed: the forward's Dropout arguments (the masks are regenerated from them)."""
    B, L, H, W = geo["B"], geo["L"], geo["H"], geo["W"]
    N = H * W
    R = B * L * N
    mode, rl, cav, T, cell = geo["mode"], geo["record_len"], geo["cav_mask"], geo["T"], geo["cell"]
    dev = sv.xs[0].device
    rows_dtype = getattr(ops, "ROWS_DTYPE", torch.bfloat16)
    common = dict(B=B, L=L, N=N, mode=mode, record_len=rl)
    geo3 = dict(B=B, L=L, N=N, record_len=rl)
    f32 = lambda: torch.zeros(B * L, C_DIM, N, dtype=torch.float32, device=dev)  # noqa: E731
    # scratch that every kernel writes before it is read, for the active agents only (padded slots are never consumed):
    # no zero fill (four 346 MB fills per step at the bench shape)
    scratch = lambda: torch.empty(B * L, C_DIM, N, dtype=torch.float32, device=dev)  # noqa: E731
    zero_b = torch.zeros(2, C_DIM, dtype=torch.float32, device=dev)
    grads = [_zero_grads(dev), _zero_grads(dev)]
    hp, dh, dz, xp = scratch(), scratch(), scratch(), scratch()
    dmk = f32() if drop_p > 0.0 else None                           # masked copy of a gradient (Dropout adjoint)
    st = torch.zeros(R, 2, dtype=torch.float32, device=dev)

    def lin_cm(a, w0, w1, bias, out, ego):
        ops.rowgemm(_lib.GEMM_LIN_CM, n_out=C_DIM, a=a, w0=w0, w1=w1, bias=bias, out=out, ego_only=ego, **common)

    head_grads = None
    if head_pack is not None:
        # out = W2 gelu(W1 x[:, 0] + b1) + b2 on the ego slot of every scene
        hg = {"w1": torch.zeros(2, C_DIM, C_DIM, device=dev), "b1": torch.zeros(2, C_DIM, device=dev),
              "w2": torch.zeros(2, C_DIM, C_DIM, device=dev), "b2": torch.zeros(2, C_DIM, device=dev)}
        d_full = f32()
        d_full.view(B, L, C_DIM, N)[:, 0] = d_out.reshape(B, C_DIM, N)
        xl = sv.xs[-1]
        lin_cm(xl, head_pack["w1_0"], head_pack["w1_1"], head_pack["b1"], hp, True)
        lin_cm(d_full, head_pack["w2T0"], head_pack["w2T1"], zero_b, dh, True)
        ops.bwd_gelu(hp, dh)
        ops.bwd_wgrad(d_full, hp, hg["w2"], ego_only=True, **common)
        ops.bwd_colsum(d_full, hg["b2"], ego_only=True, **common)
        ops.bwd_wgrad(dh, xl, hg["w1"], ego_only=True, **common)
        ops.bwd_colsum(dh, hg["b1"], ego_only=True, **common)
        dX = f32()
        lin_cm(dh, head_pack["w1T0"], head_pack["w1T1"], zero_b, dX, True)
        head_grads = hg
        del d_full
    else:
        dX = d_xlast.detach().float().reshape(B * L, C_DIM, N).clone()

    dO = torch.zeros(R, C_DIM, dtype=rows_dtype, device=dev)
    dqkv = torch.empty(5, R, C_DIM, dtype=torch.float32, device=dev)      # zero-filled before every attention backward
    dcat = torch.empty(5, R, C_DIM, dtype=rows_dtype, device=dev)         # written whole by the cast
    for s in range(len(sv.qkv) - 1, -1, -1):
        kind = s & 1
        pk, g = packs[kind], grads[kind]
        dead = sv.dead[s]
        xin, qkv, att, lse = sv.xs[s], sv.qkv[s], sv.att[s], sv.lse[s]
        dk = dict(B=B, L=L, N=N, record_len=rl, seed=seed, p=drop_p, ego_only=dead)
        # ---- recompute x' = x + [Dropout](O Wa^T + ba) and the FFN pre-activation ----
        if drop_p > 0.0:
            ops.rowgemm(_lib.GEMM_ROWS_LIN_CM, n_out=C_DIM, a=att, w0=pk["wa0"], w1=pk["wa1"], bias=pk["ba"], out=dmk,
                        ego_only=dead, **common)
            ops.dropout(dmk, xp, resid=xin, stream_id=_drop_stream(s, 0), **dk)
        else:
            ops.rowgemm(_lib.GEMM_ROWS_LIN_CM, n_out=C_DIM, a=att, w0=pk["wa0"], w1=pk["wa1"], bias=pk["ba"], resid=xin, out=xp,
                        ego_only=dead, **common)
        ops.bwd_row_stats(xp, st, ego_only=dead, **geo3)
        ops.rowgemm(_lib.GEMM_LN_LIN_CM, n_out=C_DIM, a=xp, w0=pk["w1_0"], w1=pk["w1_1"], bias=pk["b1"], out=hp,
                    ego_only=dead, ln_stats=st, **common)
        # ---- FFN: x'' = x' + [Dropout](W2 [Dropout](gelu(W1' LN(x') + b1')) + b2) ----
        dff = dX                                                     # gradient of the FFN output
        if drop_p > 0.0:
            ops.dropout(dX, dmk, stream_id=_drop_stream(s, 2), **dk)
            dff = dmk
        lin_cm(dff, pk["w2T0"], pk["w2T1"], zero_b, dh, dead)
        ops.bwd_gelu(hp, dh)                                         # hp <- gelu(hp), dh <- d pre-activation
        if drop_p > 0.0:                                             # the hidden Dropout commutes with the GELU derivative
            ops.dropout(hp, hp, stream_id=_drop_stream(s, 1), **dk)
            ops.dropout(dh, dh, stream_id=_drop_stream(s, 1), **dk)
        ops.bwd_wgrad(dff, hp, g["w2"], ego_only=dead, **common)
        ops.bwd_colsum(dff, g["b2"], ego_only=dead, **common)
        ops.bwd_wgrad(dh, xp, g["w1"], b_stats=st, ego_only=dead, **common)
        ops.bwd_colsum(dh, g["b1"], ego_only=dead, **common)
        lin_cm(dh, pk["w1T0"], pk["w1T1"], zero_b, dz, dead)
        ops.bwd_layernorm(dz, xp, st, dX, dX, ego_only=dead, **geo3)   # dX: gradient w.r.t. x'
        # ---- output projection ----
        dpr = dX                                                     # gradient of O Wa^T + ba
        if drop_p > 0.0:
            ops.dropout(dX, dmk, stream_id=_drop_stream(s, 0), **dk)
            dpr = dmk
        ops.bwd_wgrad(dpr, att, g["wa"], ego_only=dead, **common)
        ops.bwd_colsum(dpr, g["ba"], ego_only=dead, **common)
        ops.rowgemm(_lib.GEMM_LIN_ROWS, n_out=C_DIM, a=dpr, w0=pk["waT0"], w1=pk["waT1"], bias=zero_b, out=dO,
                    ego_only=dead, **common)
        # ---- attention (+ warp scatter) ----
        dqkv.zero_()
        ops.group_attn_bwd(B=B, L=L, H=H, W=W, kind=kind, mode=mode, record_len=rl, cav_mask=cav, T=T, cell=cell,
                           q=qkv[0], k=qkv[1:3], v=qkv[3:5], bk=pk["bk"], bv=pk["bv"], bias_table=pk["bias_table"],
                           o=att, d_o=dO, lse=lse, dq=dqkv[0], dk=dqkv[1:3], dv=dqkv[3:5],
                           dbk=g["bk"], dbv=g["bv"], dbias_table=g["bias_table"], ego_only=dead)
        ops.bwd_cast_colsum(dqkv, dcat, g["bcat"], B=B, L=L, N=N, mode=mode)   # bf16 copy + bias gradient of the projection
        # ---- typed LayerNorm + Q / K' / V' projection (all valid agents: they are K/V sources) ----
        st_in = sv.stats[s]                                            # saved by the forward's chain launch, else computed here
        if st_in is None:
            st_in = ops.bwd_row_stats(xin, st, **geo3)
        for p in range(5):
            ops.bwd_wgrad(dcat[p], xin, g["wcat"], b_stats=st_in, row0=p * C_DIM, **common)
        ops.bwd_dgrad_cat(dcat, pk["wcatT_0"], pk["wcatT_1"], dz, **common)
        ops.bwd_layernorm(dz, xin, st_in, dX, dX, **geo3)              # dX: gradient w.r.t. the stage input
    return dX, grads, head_grads


# ----------------------------------------------------------------------------------------------
# autograd glue
# ----------------------------------------------------------------------------------------------
class _FusionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ops, block, fusion, geo, num_iters, skip_dead, drop_p, seed, *params):
        rows_dtype = getattr(ops, "ROWS_DTYPE", torch.bfloat16)
        with torch.no_grad():
            packs = [kernel_pack_stage(fold_stage(block, k), rows_dtype) for k in ("window", "grid")]
            head_pack = kernel_pack_head(fold_head(fusion), rows_dtype) if fusion is not None else None
            xin = x.detach().float().reshape(geo["B"] * geo["L"], C_DIM, geo["H"] * geo["W"]).contiguous()
            out, xlast, sv = forward_train(ops, geo, xin, packs, head_pack, num_iters, skip_dead, drop_p, seed)
        ctx.ops, ctx.block, ctx.fusion, ctx.geo = ops, block, fusion, geo
        ctx.drop_p, ctx.seed = drop_p, seed
        ctx.sv, ctx.packs, ctx.head_pack = sv, packs, head_pack
        ctx.n_params = len(params)
        ctx.x_shape = x.shape
        B, L, H, W = geo["B"], geo["L"], geo["H"], geo["W"]
        if fusion is not None:
            return out.view(B, C_DIM, H, W)
        # padded slots pass through unchanged (like the inference path); the kernels never write them
        valid = (torch.arange(L, device=x.device)[None, :] < geo["record_len"][:, None]).view(B, L, 1, 1, 1)
        return torch.where(valid, xlast.view(B, L, C_DIM, H, W), x.detach().float())

    @staticmethod
    def backward(ctx, d_y):
        block, fusion, geo = ctx.block, ctx.fusion, ctx.geo
        with torch.no_grad():
            d_y = d_y.detach().float().contiguous()
            if fusion is not None:
                dX, grads, head_grads = backward(ctx.ops, geo, ctx.sv, ctx.packs, ctx.head_pack, d_y, None, ctx.drop_p, ctx.seed)
            else:
                dX, grads, head_grads = backward(ctx.ops, geo, ctx.sv, ctx.packs, None, None, d_y, ctx.drop_p, ctx.seed)
        ctx.sv = None                                                  # free the saved activations
        # padded slots: the kernels never touch them, so they keep 0 (head path) or the incoming gradient (block
        # path, where padded slots pass through the forward unchanged)
        dx = dX.view(ctx.x_shape)
        # pull the folded-weight gradients back to the module parameters
        owner = fusion if fusion is not None else block
        params = list(owner.parameters())
        with torch.enable_grad():
            outs, gouts = [], []
            for kind, g in zip(("window", "grid"), grads):
                F = fold_stage(block, kind)
                for k in FOLD_KEYS:
                    outs.append(F[k]); gouts.append(g[k].to(F[k].device))
            if fusion is not None:
                Fh = fold_head(fusion)
                for k in HEAD_KEYS:
                    outs.append(Fh[k]); gouts.append(head_grads[k])
            req = [p for p in params if p.requires_grad]
            keep = [i for i, o in enumerate(outs) if o.requires_grad]
            outs, gouts = [outs[i] for i in keep], [gouts[i] for i in keep]
            pg = torch.autograd.grad(outs, req, gouts, allow_unused=True) if (req and outs) else tuple(None for _ in req)
        it = iter(pg)
        gparams = [next(it) if p.requires_grad else None for p in params]
        return (dx if ctx.needs_input_grad[0] else None, None, None, None, None, None, None, None, None, *gparams)


def fusion_train(ops, block, fusion, x, pairwise_t_matrix, mode, record_len, mask, num_iters, skip_dead=True,
                 drop_p=0.0, seed=0):
    """Differentiable fusion forward (x and the module parameters).  fusion = None: block only.  drop_p > 0: train-mode
    Dropout (p of the yaml) with the Philox stream `seed`."""
    B, L, Cc, H, W = x.shape
    if tuple(pairwise_t_matrix.shape) != (B, L, L, 4, 4) or tuple(mode.shape) != (B, L) or tuple(mask.shape) != (B, L) \
            or tuple(record_len.shape) != (B,):
        raise ValueError("pairwise_t_matrix must be (B, L, L, 4, 4), mode / mask (B, L) and record_len (B,)")
    dev = x.device
    geo = {"B": B, "L": L, "H": H, "W": W,
           "T": pairwise_t_matrix.detach().to(device=dev, dtype=torch.float32).contiguous(),
           "mode": mode.detach().to(device=dev, dtype=torch.int32).contiguous(),
           "record_len": record_len.detach().to(device=dev, dtype=torch.int32).contiguous(),
           "cav_mask": mask.detach().to(device=dev, dtype=torch.int32).contiguous(),
           "cell": float(block.discrete_ratio) * float(block.downsample_rate)}
    owner = fusion if fusion is not None else block
    params = list(owner.parameters())
    return _FusionFn.apply(x, ops, block, fusion, geo, num_iters, skip_dead, float(drop_p), int(seed), *params)
