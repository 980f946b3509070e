"""TEST INFRASTRUCTURE ONLY -- CPU emulation of the CUDA path's order of operations.

Same algorithm as oracle/hmvit_oracle.py (and therefore as the reference), but evaluated in the
RESTRUCTURED order the sm_100a kernels use (DESIGN.md "exact restructurings"):
  * project-then-warp: K/V are projected once per agent and the bilinear warp gathers the projected
    rows (the warp is linear over space, biases are added after the gather);
  * edge-type weights folded into the projections:  K' = blockdiag_h(relation_att[e]) W_k,
    V' = blockdiag_h(relation_msg[e]^T) W_v with e = type(ego)*2 + type(source);
  * the softmax scale is folded into W_q;
and with an optional rounding function `rnd` applied at exactly the points where the kernels round to
bf16 (GEMM operands, stored Q/K/V, gathered K/V, softmax probabilities, attention output, FFN hidden).
With rnd = identity this must agree with the oracle to fp32 round-off (tests/test_emul.py); with
rnd = bf16 it is the tight checker for the kernels and measures the error budget of bf16 operands
against the fp32 reference.
"""
from __future__ import annotations

from typing import Callable, Dict

import torch
import torch.nn.functional as F

from . import hmvit_oracle as O

Tensor = torch.Tensor


def bf16_round(t: Tensor) -> Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def identity(t: Tensor) -> Tensor:
    return t


SITES = ("xhat", "w", "qkv", "kvg", "p", "o", "hid")


class Rounding:
    """Per-site rounding: Rounding(bf16_round) rounds everywhere, Rounding(bf16_round, only=("w",))
    rounds the weights only, Rounding() is exact."""

    def __init__(self, fn: Callable[[Tensor], Tensor] = identity, only=None, skip=()):
        self.fn = fn
        self.sites = set(SITES if only is None else only) - set(skip)

    def __call__(self, t: Tensor, site: str) -> Tensor:
        assert site in SITES, site
        return self.fn(t) if site in self.sites else t


def pack_attention_weights(P: Dict[str, Tensor], pfx: str, dim_head: int = 32):
    """Folded projection weights of one attention module.

    Returns dict with
      wq[t] (C,C), bq[t] (C,)                 scale folded in
      wk[te][t], bk[te][t], wv[te][t], bv[te][t]   edge type e = te*2+t folded in
      wa[t], ba[t]
    """
    C = P[f"{pfx}.q_linears.0.weight"].shape[0]
    h, d = C // dim_head, dim_head
    scale = d ** -0.5
    att, msg = P[f"{pfx}.relation_att"], P[f"{pfx}.relation_msg"]
    out = {"wq": {}, "bq": {}, "wk": {0: {}, 1: {}}, "bk": {0: {}, 1: {}}, "wv": {0: {}, 1: {}},
           "bv": {0: {}, 1: {}}, "wa": {}, "ba": {}}
    for t in (0, 1):
        out["wq"][t] = P[f"{pfx}.q_linears.{t}.weight"] * scale
        out["bq"][t] = P[f"{pfx}.q_linears.{t}.bias"] * scale
        out["wa"][t] = P[f"{pfx}.a_linears.{t}.0.weight"]
        out["ba"][t] = P[f"{pfx}.a_linears.{t}.0.bias"]
        wk = P[f"{pfx}.k_linears.{t}.weight"].view(h, d, C)
        bk = P[f"{pfx}.k_linears.{t}.bias"].view(h, d)
        wv = P[f"{pfx}.v_linears.{t}.weight"].view(h, d, C)
        bv = P[f"{pfx}.v_linears.{t}.bias"].view(h, d)
        for te in (0, 1):
            e = te * 2 + t
            # k'[h,p] = sum_q att[e,h,p,q] k[h,q]      v'[h,q] = sum_p msg[e,h,p,q] v[h,p]
            out["wk"][te][t] = torch.einsum("hpq,hqc->hpc", att[e], wk).reshape(C, C)
            out["bk"][te][t] = torch.einsum("hpq,hq->hp", att[e], bk).reshape(C)
            out["wv"][te][t] = torch.einsum("hpq,hpc->hqc", msg[e], wv).reshape(C, C)
            out["bv"][te][t] = torch.einsum("hpq,hp->hq", msg[e], bv).reshape(C)
    return out


def fusion_stage(x: Tensor, T: Tensor, mode: Tensor, record_len: Tensor, cav_mask: Tensor,
                 P: Dict[str, Tensor], pfx: str, kind: str, cfg: dict,
                 rnd: "Rounding" = Rounding(), ego_only: bool = False) -> Tensor:
    """x (B, L, H, W, C) fp32 residual stream -> same.  Only valid slots l < record_len[b] are
    updated (padded / out-of-scene slots are left untouched: they never influence valid ones).
    ego_only: compute the update for agent 0 only (exact dead-query elimination for the last stage)."""
    B, L, H, W, C = x.shape
    w, dh = cfg["window_size"], cfg["dim_head"]
    h = C // dh
    dr, ds = cfg["spatial_transform"]["voxel_size"][0], cfg["spatial_transform"]["downsample_rate"]
    table = O.group_token_table(H, W, w, kind)
    G, S = table.shape
    apfx = f"{pfx}.{kind}_attention"
    pk = pack_attention_weights(P, apfx, dh)
    rel = O.relative_position_index(w)
    bias = P[f"{apfx}.relative_position_bias_table.weight"][rel].permute(2, 0, 1)   # (h, S, S)
    out = x.clone()
    for b in range(B):
        n = int(record_len[b])
        types = [int(v) for v in mode[b, :n]]
        xh = [rnd(F.layer_norm(x[b, l], (C,), P[f"{pfx}.{kind}_norm.net.{types[l]}.weight"],
                               P[f"{pfx}.{kind}_norm.net.{types[l]}.bias"], 1e-5), "xhat").reshape(H * W, C)
              for l in range(n)]
        ego_types = sorted(set(types[:1] if ego_only else types))
        Q = [rnd(F.linear(xh[l], rnd(pk["wq"][types[l]], "w"), pk["bq"][types[l]]), "qkv") for l in range(n)]
        K = {te: [rnd(F.linear(xh[l], rnd(pk["wk"][te][types[l]], "w")), "qkv") for l in range(n)] for te in ego_types}
        V = {te: [rnd(F.linear(xh[l], rnd(pk["wv"][te][types[l]], "w")), "qkv") for l in range(n)] for te in ego_types}
        for i in range(1 if ego_only else n):
            te = types[i]
            sx, sy = O.source_coords(T[b, :n, i], H, W, dr, ds)
            rx, ry = torch.round(sx), torch.round(sy)
            km = ((rx >= 0) & (rx <= W - 1) & (ry >= 0) & (ry <= H - 1)) & (cav_mask[b, :n].view(n, 1, 1) != 0)
            km = km.reshape(n, H * W)[:, table]                                        # (n, G, S)
            Kw = O.warp_bilinear_nhwc(torch.stack(K[te]).view(n, H, W, C), sx, sy).reshape(n, H * W, C)
            Vw = O.warp_bilinear_nhwc(torch.stack(V[te]).view(n, H, W, C), sx, sy).reshape(n, H * W, C)
            bk = torch.stack([pk["bk"][te][types[j]] for j in range(n)]).view(n, 1, C)
            bv = torch.stack([pk["bv"][te][types[j]] for j in range(n)]).view(n, 1, C)
            Kg = rnd(Kw + bk, "kvg")[:, table].view(n, G, S, h, dh)
            Vg = rnd(Vw + bv, "kvg")[:, table].view(n, G, S, h, dh)
            q = Q[i][table].view(G, S, h, dh)
            logits = torch.einsum("gshd,jgkhd->ghsjk", q, Kg) + bias[None, :, :, None, :]
            logits = logits.masked_fill(~km.permute(1, 0, 2).reshape(G, 1, 1, n, S), float("-inf"))
            logits = logits.reshape(G, h, S, n * S)
            m = logits.max(-1, keepdim=True).values
            p = torch.exp(logits - m)
            o = torch.einsum("ghsjk,jgkhd->gshd", rnd(p, "p").view(G, h, S, n, S), Vg) / p.sum(-1).permute(0, 2, 1)[..., None]
            o = rnd(o.reshape(G * S, C), "o")
            y = F.linear(o, rnd(pk["wa"][te], "w"), pk["ba"][te])
            x1 = x[b, i].reshape(H * W, C).clone()
            x1[table.reshape(-1)] += y
            # FFN with pre-norm and residual
            fp = f"{pfx}.{kind}_ffd"
            xn = rnd(F.layer_norm(x1, (C,), P[f"{fp}.norm.net.{te}.weight"], P[f"{fp}.norm.net.{te}.bias"], 1e-5), "xhat")
            hid = rnd(F.gelu(F.linear(xn, rnd(P[f"{fp}.fn.net.{te}.0.weight"], "w"), P[f"{fp}.fn.net.{te}.0.bias"])), "hid")
            x2 = x1 + F.linear(hid, rnd(P[f"{fp}.fn.net.{te}.3.weight"], "w"), P[f"{fp}.fn.net.{te}.3.bias"])
            out[b, i] = x2.view(H, W, C)
    return out


def hetero_fusion(x: Tensor, T: Tensor, mode: Tensor, record_len: Tensor, cav_mask: Tensor,
                  P: Dict[str, Tensor], config: dict, rnd: "Rounding" = Rounding(),
                  skip_dead: bool = True) -> Tensor:
    """(B, L, C, H, W) -> (B, C, H, W); valid slots only."""
    mode = mode.to(torch.int64)
    cfg = config["hetero_fusion_block"]
    y = x.permute(0, 1, 3, 4, 2).contiguous()
    n_it = config["num_iters"]
    for it in range(n_it):
        y = fusion_stage(y, T, mode, record_len, cav_mask, P, "hetero_fusion_block", "window", cfg, rnd)
        y = fusion_stage(y, T, mode, record_len, cav_mask, P, "hetero_fusion_block", "grid", cfg, rnd,
                         ego_only=skip_dead and it == n_it - 1)
    B, L, H, W, C = y.shape
    out = y.new_empty(B, H, W, C)
    for b in range(B):
        te = int(mode[b, 0])
        hid = rnd(F.gelu(F.linear(rnd(y[b, 0], "xhat"), rnd(P[f"mlp_head.net.{te}.0.weight"], "w"), P[f"mlp_head.net.{te}.0.bias"])), "hid")
        out[b] = F.linear(hid, rnd(P[f"mlp_head.net.{te}.3.weight"], "w"), P[f"mlp_head.net.{te}.3.bias"])
    return out.permute(0, 3, 1, 2).contiguous()
