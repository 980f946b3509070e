"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the HM-ViT fusion forward.

A plain PyTorch-CPU (fp32) restatement of the reference algorithm, written
from the reference's semantics and following its order of operations
(LayerNorm -> pairwise warp -> per-ego attention -> residual -> FFN, window
stage then grid stage).  It is the checker for the CUDA path and the timed
"port" CPU baseline in bench.py; it is never imported by the product package.

Parity pin: the reference holds no golden vectors for this path
(/root/reference/test covers box/pcd utils only), so the oracle is pinned to
OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports the
unmodified reference from /root/reference in the build container, runs it on
seeded inputs and commits the outputs under tests/golden/;
tests/test_oracle_golden.py checks this file against those vectors.

Reference files followed (all under /root/reference/opencood/models/):
  bevformer_point_pillar_hetero.py:22-49      HeteroFusion
  sub_modules/hetero_fusion.py:16-277         HeteroAttention
  sub_modules/hetero_fusion.py:279-474        HeteroFusionBlock
  base_transformer.py:121-192                 HeteroLayerNorm / HeteroFeedForward / PreNormResidual
  sub_modules/spatial_transformation.py:16-44 SpatialTransformation
  sub_modules/torch_transformation_utils.py:11-134,216-355  warp / ROI mask
  sub_modules/fuse_utils.py:8-61              regroup

All tensors are torch CPU tensors; `P` is a dict with the reference's
state_dict keys (prefix-less, i.e. the keys of HeteroFusion.state_dict()).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------
# index maps (integer, must be bit exact)
# --------------------------------------------------------------------------
def partition_index(H: int, W: int, w: int, kind: str) -> Tuple[Tensor, Tensor]:
    """token (r, c) -> (group id, slot id) for the two partitions.

    window: 'b m d (x w1) (y w2) -> b m x y w1 w2 d'  hetero_fusion.py:384-389
        group = (r // w, c // w), slot = (r % w, c % w)
    grid:   'b m d (w1 x) (w2 y) -> b m x y w1 w2 d'  hetero_fusion.py:427-431
        group = (r % (H/w), c % (W/w)), slot = (r // (H/w), c // (W/w))
    Returns int64 tensors of shape (H, W): group in [0, (H/w)*(W/w)), slot in [0, w*w).
    """
    if H % w or W % w:
        raise ValueError(f"H={H}, W={W} must be divisible by window {w}")
    X, Y = H // w, W // w
    r = torch.arange(H).view(H, 1).expand(H, W)
    c = torch.arange(W).view(1, W).expand(H, W)
    if kind == "window":
        gx, gy, s1, s2 = r // w, c // w, r % w, c % w
    elif kind == "grid":
        gx, gy, s1, s2 = r % X, c % Y, r // X, c // Y
    else:
        raise ValueError(kind)
    return (gx * Y + gy).contiguous(), (s1 * w + s2).contiguous()


def group_token_table(H: int, W: int, w: int, kind: str) -> Tensor:
    """(G, S) int64 table: flat token index r*W+c of slot s in group g."""
    group, slot = partition_index(H, W, w, kind)
    G, S = (H // w) * (W // w), w * w
    table = torch.empty(G, S, dtype=torch.int64)
    table[group.reshape(-1), slot.reshape(-1)] = torch.arange(H * W)
    return table


def relative_position_index(w: int) -> Tensor:
    """hetero_fusion.py:82-109 -- (w*w, w*w) int64, value (dr+w-1)*(2w-1) + (dc+w-1)."""
    r = torch.arange(w).view(w, 1).expand(w, w).reshape(-1)
    c = torch.arange(w).view(1, w).expand(w, w).reshape(-1)
    dr = r[:, None] - r[None, :] + (w - 1)
    dc = c[:, None] - c[None, :] + (w - 1)
    return dr * (2 * w - 1) + dc


# --------------------------------------------------------------------------
# regroup (fuse_utils.py:8-61)
# --------------------------------------------------------------------------
def regroup(dense_feature: Tensor, record_len: Tensor, max_len: int) -> Tuple[Tensor, Tensor]:
    """(sum L_b, C, H, W) -> zero padded (B, L, C, H, W) and int64 (B, L) mask."""
    lens = [int(v) for v in record_len.tolist()]
    _, C, H, W = dense_feature.shape
    out = dense_feature.new_zeros(len(lens), max_len, C, H, W)
    mask = torch.zeros(len(lens), max_len, dtype=torch.int64)
    start = 0
    for b, n in enumerate(lens):
        out[b, :n] = dense_feature[start:start + n]
        mask[b, :n] = 1
        start += n
    return out, mask


# --------------------------------------------------------------------------
# warp geometry (torch_transformation_utils.py:108-134, 254-297, 317-355)
# --------------------------------------------------------------------------
def source_coords(T: Tensor, H: int, W: int, discrete_ratio: float, downsample_rate: float) -> Tuple[Tensor, Tensor]:
    """Source pixel coordinates sampled by each output pixel.

    T: (..., 4, 4) transform "source agent -> target agent".  The reference
    keeps rows {0,1} x cols {0,1,3}, divides the translation by
    discrete_ratio*downsample_rate (:129-134), builds  dst = A (src - c) + c + t
    with c = (W/2, H/2) (:254-297), inverts it and samples with
    align_corners=True (pixel centre = integer coordinate).  Closed form:
        src = A^-1 ((u, v) - c - t) + c
    Returns float64 (..., H, W) tensors (sx, sy).
    """
    T64 = T.to(torch.float64)
    A = T64[..., :2, :2]
    t = T64[..., :2, 3] / (float(discrete_ratio) * float(downsample_rate))
    Ainv = torch.linalg.inv(A)
    u = torch.arange(W, dtype=torch.float64).view(1, W)
    v = torch.arange(H, dtype=torch.float64).view(H, 1)
    du = u - W / 2.0 - t[..., 0, None, None]
    dv = v - H / 2.0 - t[..., 1, None, None]
    sx = Ainv[..., 0, 0, None, None] * du + Ainv[..., 0, 1, None, None] * dv + W / 2.0
    sy = Ainv[..., 1, 0, None, None] * du + Ainv[..., 1, 1, None, None] * dv + H / 2.0
    return sx, sy


def warp_bilinear_nhwc(x: Tensor, sx: Tensor, sy: Tensor) -> Tensor:
    """Bilinear gather with zero padding (F.grid_sample semantics, align_corners=True).

    x: (n, H, W, C) float32; sx, sy: (n, H, W) float64.  Returns (n, H, W, C).
    """
    n, H, W, C = x.shape
    x0 = torch.floor(sx)
    y0 = torch.floor(sy)
    wx1 = (sx - x0)
    wy1 = (sy - y0)
    wx0 = 1.0 - wx1
    wy0 = 1.0 - wy1
    out = x.new_zeros(n, H * W, C)
    xf = x.reshape(n, H * W, C)
    for dx, dy, wgt in ((0, 0, wx0 * wy0), (1, 0, wx1 * wy0), (0, 1, wx0 * wy1), (1, 1, wx1 * wy1)):
        xi = (x0 + dx).to(torch.int64)
        yi = (y0 + dy).to(torch.int64)
        ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
        idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).reshape(n, H * W, 1).expand(n, H * W, C)
        wgt = (wgt * ok).to(torch.float32).reshape(n, H * W, 1)
        out += torch.gather(xf, 1, idx) * wgt
    return out.reshape(n, H, W, C)


def spatial_transformation(x: Tensor, T: Tensor, discrete_ratio: float, downsample_rate: float) -> Tensor:
    """spatial_transformation.py:16-44 -- x (B, L, C, H, W), T (B, L, 4, 4) -> (B, L, C, H, W)."""
    B, L, C, H, W = x.shape
    sx, sy = source_coords(T, H, W, discrete_ratio, downsample_rate)
    y = warp_bilinear_nhwc(x.permute(0, 1, 3, 4, 2).reshape(B * L, H, W, C),
                           sx.reshape(B * L, H, W), sy.reshape(B * L, H, W))
    return y.reshape(B, L, H, W, C).permute(0, 1, 4, 2, 3).contiguous()


def roi_mask(T: Tensor, H: int, W: int, discrete_ratio: float, downsample_rate: float) -> Tensor:
    """Nearest warp of a tensor of ones (torch_transformation_utils.py:77-105): 1.0 where the
    rounded (half-to-even) source coordinate lies inside the map.  Returns float32 (..., H, W)."""
    sx, sy = source_coords(T, H, W, discrete_ratio, downsample_rate)
    rx, ry = torch.round(sx), torch.round(sy)          # torch.round == rint (half to even)
    ok = (rx >= 0) & (rx <= W - 1) & (ry >= 0) & (ry <= H - 1)
    return ok.to(torch.float32)


def roi_and_cav_mask(shape, cav_mask: Tensor, T: Tensor, discrete_ratio: float, downsample_rate: float) -> Tensor:
    """get_roi_and_cav_mask (torch_transformation_utils.py:11-49) -> float32 (B, H, W, 1, L)."""
    B, L, H, W, _ = shape
    roi = roi_mask(T, H, W, discrete_ratio, downsample_rate)             # (B, L, H, W)
    com = roi * cav_mask.to(torch.float32).view(B, L, 1, 1)
    return com.permute(0, 2, 3, 1).unsqueeze(3).contiguous()


# --------------------------------------------------------------------------
# typed LayerNorm / FFN (base_transformer.py:138-192)
# --------------------------------------------------------------------------
def hetero_layer_norm(x: Tensor, mode: Tensor, P: Dict[str, Tensor], pfx: str, eps: float = 1e-5) -> Tensor:
    """x (B, L, ..., C); each (b, l) slice uses LayerNorm `pfx.net.{mode[b,l]}`."""
    out = torch.empty_like(x)
    C = x.shape[-1]
    for t in (0, 1):
        sel = mode == t
        if sel.any():
            out[sel] = F.layer_norm(x[sel], (C,), P[f"{pfx}.net.{t}.weight"], P[f"{pfx}.net.{t}.bias"], eps)
    return out


def hetero_ffn(x: Tensor, mode: Tensor, P: Dict[str, Tensor], pfx: str, drop=None) -> Tensor:
    """Linear -> GELU(erf) -> Dropout -> Linear -> Dropout per type (base_transformer.py:180-192).  Dropout is the
    identity in eval; training parity passes the scaled keep masks drop = {"hid": ..., "ffn": ...} (same shape as x)."""
    out = None
    for t in (0, 1):
        sel = mode == t
        if sel.any():
            h = F.gelu(F.linear(x[sel], P[f"{pfx}.net.{t}.0.weight"], P[f"{pfx}.net.{t}.0.bias"]))
            if drop is not None:
                h = h * drop["hid"][sel]
            y = F.linear(h, P[f"{pfx}.net.{t}.3.weight"], P[f"{pfx}.net.{t}.3.bias"])
            if drop is not None:
                y = y * drop["ffn"][sel]
            if out is None:
                out = x.new_empty(*x.shape[:-1], y.shape[-1])
            out[sel] = y
    return out


# --------------------------------------------------------------------------
# attention for one (scene, ego) (hetero_fusion.py:187-277)
# --------------------------------------------------------------------------
def hetero_attention_ego(xw: Tensor, types: Tensor, ego: int, keymask: Tensor, P: Dict[str, Tensor],
                         pfx: str, dim_head: int = 32, window: int = 8) -> Tensor:
    """xw (Ls, G, S, C): normalised features of every source already warped into the ego frame and
    partitioned into G groups of S=window^2 slots; types (Ls,) int; keymask (Ls, G, S) {0,1}.
    Returns the attention output for the ego's tokens, (G, S, C) (after the typed output Linear).
    """
    Ls, G, S, C = xw.shape
    h, d = C // dim_head, dim_head
    te = int(types[ego])
    q = F.linear(xw[ego], P[f"{pfx}.q_linears.{te}.weight"], P[f"{pfx}.q_linears.{te}.bias"])
    q = q.view(G, S, h, d) * (d ** -0.5)
    rel = relative_position_index(window)
    bias = P[f"{pfx}.relative_position_bias_table.weight"][rel]          # (S, S, h)
    logits = xw.new_empty(G, h, S, Ls, S)
    vmsg = xw.new_empty(Ls, G, S, h, d)
    for j in range(Ls):
        tj = int(types[j])
        e = te * 2 + tj                                                  # :154-155
        k = F.linear(xw[j], P[f"{pfx}.k_linears.{tj}.weight"], P[f"{pfx}.k_linears.{tj}.bias"]).view(G, S, h, d)
        v = F.linear(xw[j], P[f"{pfx}.v_linears.{tj}.weight"], P[f"{pfx}.v_linears.{tj}.bias"]).view(G, S, h, d)
        qa = torch.einsum("gshp,hpq->gshq", q, P[f"{pfx}.relation_att"][e])
        logits[:, :, :, j, :] = torch.einsum("gshq,gkhq->ghsk", qa, k) + bias.permute(2, 0, 1)[None]
        vmsg[j] = torch.einsum("gkhp,hpq->gkhq", v, P[f"{pfx}.relation_msg"][e])
    km = keymask.permute(1, 0, 2).reshape(G, 1, 1, Ls, S)
    logits = logits.masked_fill(km == 0, float("-inf"))
    attn = torch.softmax(logits.reshape(G, h, S, Ls * S), dim=-1).view(G, h, S, Ls, S)
    out = torch.einsum("ghsjk,jgkhq->gshq", attn, vmsg).reshape(G, S, C)
    return F.linear(out, P[f"{pfx}.a_linears.{te}.0.weight"], P[f"{pfx}.a_linears.{te}.0.bias"])


# --------------------------------------------------------------------------
# one stage (window or grid) of the fusion block (hetero_fusion.py:363-444)
# --------------------------------------------------------------------------
def fusion_stage(x: Tensor, T: Tensor, mode: Tensor, record_len: Tensor, cav_mask: Tensor,
                 P: Dict[str, Tensor], pfx: str, kind: str, cfg: dict, drop=None) -> Tensor:
    """x (B, L, H, W, C) channels-last residual stream -> same shape.  drop: optional scaled keep masks of the three
    train-mode Dropout sites of the stage, {"att", "hid", "ffn"}, each (B, L, H, W, C) (hetero_fusion.py:66,
    base_transformer.py:186-190); None = eval."""
    B, L, H, W, C = x.shape
    w = cfg["window_size"]
    dr, ds = cfg["spatial_transform"]["voxel_size"][0], cfg["spatial_transform"]["downsample_rate"]
    table = group_token_table(H, W, w, kind)                                 # (G, S)
    G, S = table.shape
    xn = hetero_layer_norm(x, mode, P, f"{pfx}.{kind}_norm")
    Lv = int(record_len.max())
    upd = torch.zeros_like(x)
    for b in range(B):
        for i in range(Lv):                                                  # ego loop :376
            Tbi = T[b, :Lv, i]                                               # all sources -> target i (:345)
            sx, sy = source_coords(Tbi, H, W, dr, ds)
            xw = warp_bilinear_nhwc(xn[b, :Lv], sx, sy)                      # (Lv, H, W, C)
            rx, ry = torch.round(sx), torch.round(sy)
            roi = ((rx >= 0) & (rx <= W - 1) & (ry >= 0) & (ry <= H - 1)).to(torch.float32)
            km = roi * cav_mask[b, :Lv].to(torch.float32).view(Lv, 1, 1)     # (Lv, H, W)
            xw = xw.reshape(Lv, H * W, C)[:, table]                          # (Lv, G, S, C)
            km = km.reshape(Lv, H * W)[:, table]
            y = hetero_attention_ego(xw, mode[b, :Lv], i, km, P, f"{pfx}.{kind}_attention",
                                     cfg["dim_head"], w)
            upd[b, i].view(H * W, C)[table.reshape(-1)] = y.reshape(G * S, C)
    if drop is not None:
        upd = upd * drop["att"]                                              # Dropout behind a_linears (:66)
    x = x + upd                                                              # :399
    xn2 = hetero_layer_norm(x, mode, P, f"{pfx}.{kind}_ffd.norm")
    return x + hetero_ffn(xn2, mode, P, f"{pfx}.{kind}_ffd.fn", drop)        # :401


def fusion_block(x: Tensor, T: Tensor, mode: Tensor, record_len: Tensor, cav_mask: Tensor,
                 P: Dict[str, Tensor], cfg: dict, pfx: str = "hetero_fusion_block", drop_masks=None) -> Tensor:
    """HeteroFusionBlock.forward, architect_mode == 'sequential' (hetero_fusion.py:446-458).
    x (B, L, C, H, W) -> (B, L, C, H, W)."""
    if cfg.get("architect_mode", "sequential") != "sequential":
        raise ValueError(f"{cfg.get('architect_mode')} not implemented")
    y = x.permute(0, 1, 3, 4, 2).contiguous()
    y = fusion_stage(y, T, mode, record_len, cav_mask, P, pfx, "window", cfg, drop_masks[0] if drop_masks is not None else None)
    y = fusion_stage(y, T, mode, record_len, cav_mask, P, pfx, "grid", cfg, drop_masks[1] if drop_masks is not None else None)
    return y.permute(0, 1, 4, 2, 3).contiguous()


def hetero_fusion(x: Tensor, T: Tensor, mode: Tensor, record_len: Tensor, cav_mask: Tensor,
                  P: Dict[str, Tensor], config: dict, drop_masks=None) -> Tensor:
    """HeteroFusion.forward (bevformer_point_pillar_hetero.py:39-49): num_iters x the same block,
    ego slice, typed FFN head (no norm, no residual; its Dropout has p = 0).  Returns (B, C, H, W).
    drop_masks: optional list (one entry per stage, 2 * num_iters) of train-mode Dropout masks, see fusion_stage."""
    mode = mode.to(torch.int64)
    cfg = config["hetero_fusion_block"]
    y = x.permute(0, 1, 3, 4, 2).contiguous()
    for it in range(config["num_iters"]):
        dw = drop_masks[2 * it] if drop_masks is not None else None
        dg = drop_masks[2 * it + 1] if drop_masks is not None else None
        y = fusion_stage(y, T, mode, record_len, cav_mask, P, "hetero_fusion_block", "window", cfg, dw)
        y = fusion_stage(y, T, mode, record_len, cav_mask, P, "hetero_fusion_block", "grid", cfg, dg)
    ego = y[:, :1]                                                           # (B, 1, H, W, C)
    out = hetero_ffn(ego, mode[:, :1], P, "mlp_head")
    return out[:, 0].permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------
# synthetic parameters / inputs shared by tests, golden generation and bench
# --------------------------------------------------------------------------
def default_config(input_dim: int = 256, window: int = 8, dim_head: int = 32, num_iters: int = 2,
                   agent_size: int = 5) -> dict:
    st = {"voxel_size": [0.4, 0.4, 4], "downsample_rate": 4, "use_roi_mask": True}
    return {"num_iters": num_iters, "spatial_transform": st,
            "hetero_fusion_block": {"spatial_transform": st, "architect_mode": "sequential",
                                    "input_dim": input_dim, "mlp_dim": input_dim, "agent_size": agent_size,
                                    "window_size": window, "dim_head": dim_head, "drop_out": 0.1, "mask": True}}


def state_dict_spec(config: dict):
    """(key, shape) list of HeteroFusion.state_dict() in the reference's registration order."""
    cfg = config["hetero_fusion_block"]
    C, M, w, d = cfg["input_dim"], cfg["mlp_dim"], cfg["window_size"], cfg["dim_head"]
    h = C // d
    spec = []

    def ln(p):
        for t in (0, 1):
            spec.extend([(f"{p}.net.{t}.weight", (C,)), (f"{p}.net.{t}.bias", (C,))])

    def ffn(p, din, dh, dout):
        for t in (0, 1):
            spec.extend([(f"{p}.net.{t}.0.weight", (dh, din)), (f"{p}.net.{t}.0.bias", (dh,)),
                         (f"{p}.net.{t}.3.weight", (dout, dh)), (f"{p}.net.{t}.3.bias", (dout,))])

    def attn(p):
        spec.extend([(f"{p}.relation_att", (4, h, d, d)), (f"{p}.relation_msg", (4, h, d, d)),
                     (f"{p}.relative_position_index", (w * w, w * w))])
        for name in ("k_linears", "q_linears", "v_linears"):
            for t in (0, 1):
                spec.extend([(f"{p}.{name}.{t}.weight", (C, C)), (f"{p}.{name}.{t}.bias", (C,))])
        for t in (0, 1):
            spec.extend([(f"{p}.a_linears.{t}.0.weight", (C, C)), (f"{p}.a_linears.{t}.0.bias", (C,))])
        spec.append((f"{p}.relative_position_bias_table.weight", ((2 * w - 1) ** 2, h)))

    b = "hetero_fusion_block"
    ln(f"{b}.window_norm"); attn(f"{b}.window_attention"); ln(f"{b}.window_ffd.norm"); ffn(f"{b}.window_ffd.fn", C, M, C)
    ln(f"{b}.grid_norm"); attn(f"{b}.grid_attention"); ln(f"{b}.grid_ffd.norm"); ffn(f"{b}.grid_ffd.fn", C, M, C)
    ffn(f"{b}.aggregate_fc", M * 3, M, M)
    ffn("mlp_head", C, C, C)
    return spec


def synth_state_dict(config: dict, seed: int = 0) -> Dict[str, Tensor]:
    """Deterministic synthetic parameters (independent of nn.Module construction order):
    LN weight 1+0.1n, biases 0.1n, Linear weight n/sqrt(fan_in), relation_* xavier-like, table N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    w = config["hetero_fusion_block"]["window_size"]
    for key, shape in state_dict_spec(config):
        if key.endswith("relative_position_index"):
            P[key] = relative_position_index(w)
        elif "relation_" in key:
            fan = shape[1] * shape[2] * shape[3]
            P[key] = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(6.0 / (fan + shape[0] * shape[2] * shape[3]))
        elif key.endswith("bias_table.weight"):
            P[key] = torch.randn(shape, generator=g)
        elif "norm" in key and key.endswith("weight"):
            P[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            P[key] = 0.1 * torch.randn(shape, generator=g)
        else:
            P[key] = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
    return P


def synth_inputs(B: int, L: int, C: int, H: int, W: int, record_len, seed: int, camera_p: float = 0.5,
                 mode=None, tx: float = 50.0, ty: float = 15.0):
    """Synthetic scene batch: x ~ N(0,1) regrouped with zero padding; ego pose identity,
    collaborators yaw~U(-pi,pi), tx~U(-tx,tx) m, ty~U(-ty,ty) m; T[b,i,j] = P_j^-1 P_i, identity on
    the diagonal and on padded slots; mode ~ Bernoulli(camera_p) (0=camera, 1=lidar), padding 0."""
    g = torch.Generator().manual_seed(seed)
    record_len = torch.as_tensor(record_len, dtype=torch.int64)
    assert record_len.numel() == B and int(record_len.max()) <= L and int(record_len.min()) >= 1
    dense = torch.randn(int(record_len.sum()), C, H, W, generator=g)
    x, mask = regroup(dense, record_len, L)
    yaw = (torch.rand(B, L, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
    px = (torch.rand(B, L, generator=g, dtype=torch.float64) * 2 - 1) * tx
    py = (torch.rand(B, L, generator=g, dtype=torch.float64) * 2 - 1) * ty
    yaw[:, 0], px[:, 0], py[:, 0] = 0, 0, 0
    pose = torch.eye(4, dtype=torch.float64).repeat(B, L, 1, 1)
    pose[..., 0, 0], pose[..., 0, 1] = torch.cos(yaw), -torch.sin(yaw)
    pose[..., 1, 0], pose[..., 1, 1] = torch.sin(yaw), torch.cos(yaw)
    pose[..., 0, 3], pose[..., 1, 3] = px, py
    T = torch.eye(4, dtype=torch.float64).repeat(B, L, L, 1, 1)
    for b in range(B):
        n = int(record_len[b])
        for i in range(n):
            for j in range(n):
                if i != j:
                    T[b, i, j] = torch.linalg.inv(pose[b, j]) @ pose[b, i]
    if mode is None:
        mode = (torch.rand(B, L, generator=g) >= camera_p).to(torch.int32)
    else:
        mode = torch.as_tensor(mode, dtype=torch.int32).clone()
    mode = mode * mask.to(torch.int32)                                      # padded slots are camera (0)
    return x, T.to(torch.float32), mode, record_len, mask


# --------------------------------------------------------------------------
# detection heads after the fusion (test infrastructure for the "logits" parity row)
# --------------------------------------------------------------------------
def decoder_state_spec(input_dim: int = 256, num_layer: int = 2, num_ch_dec=(256, 256), anchor_number: int = 2):
    """(key, shape) list of HeteroDecoder.state_dict() (hetero_decoder.py:27-40, naive_decoder.py:27-54):
    per modality 2 x (conv3x3-BN-ReLU, conv3x3-BN-ReLU) + 1x1 cls / reg heads."""
    spec = []
    for mod in ("camera", "lidar"):
        idx = 0
        for i in range(num_layer - 1, -1, -1):
            cin = input_dim if i == num_layer - 1 else num_ch_dec[i + 1]
            cout = num_ch_dec[i]
            for c_in in (cin, cout):
                spec.append((f"{mod}_decoder.decoder.{idx}.weight", (cout, c_in, 3, 3)))
                spec.append((f"{mod}_decoder.decoder.{idx}.bias", (cout,)))
                for nm, shp in (("weight", (cout,)), ("bias", (cout,)), ("running_mean", (cout,)),
                                ("running_var", (cout,)), ("num_batches_tracked", ())):
                    spec.append((f"{mod}_decoder.decoder.{idx + 1}.{nm}", shp))
                idx += 3
    for mod in ("camera", "lidar"):
        spec.append((f"{mod}_cls_head.weight", (anchor_number, num_ch_dec[0], 1, 1)))
        spec.append((f"{mod}_cls_head.bias", (anchor_number,)))
        spec.append((f"{mod}_reg_head.weight", (7 * anchor_number, num_ch_dec[0], 1, 1)))
        spec.append((f"{mod}_reg_head.bias", (7 * anchor_number,)))
    return spec


def synth_decoder_state_dict(seed: int = 0, **kw) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    P = {}
    for key, shape in decoder_state_spec(**kw):
        if key.endswith("num_batches_tracked"):
            P[key] = torch.tensor(0, dtype=torch.int64)
        elif key.endswith("running_var"):
            P[key] = 0.5 + torch.rand(shape, generator=g)
        elif key.endswith("running_mean") or key.endswith("bias"):
            P[key] = 0.1 * torch.randn(shape, generator=g)
        elif len(shape) == 1:                       # BN weight
            P[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                       # conv weight
            fan_in = shape[1] * shape[2] * shape[3]
            P[key] = torch.randn(shape, generator=g) / math.sqrt(fan_in)
    return P


def hetero_decoder(x: Tensor, ego_mode: Tensor, P: Dict[str, Tensor], num_layer: int = 2):
    """HeteroDecoder.forward with use_upsample=False, eval-mode BatchNorm (hetero_decoder.py:42-74,
    naive_decoder.py:63-92): x (B, C, H, W) fused ego feature, ego_mode (B,) -> psm (B, A, H, W), rm (B, 7A, H, W)."""
    psm, rm = [], []
    for b in range(x.shape[0]):
        mod = "lidar" if int(ego_mode[b]) == 1 else "camera"
        y = x[b:b + 1]
        for blk in range(2 * num_layer):
            i = 3 * blk
            y = F.conv2d(y, P[f"{mod}_decoder.decoder.{i}.weight"], P[f"{mod}_decoder.decoder.{i}.bias"], padding=1)
            y = F.batch_norm(y, P[f"{mod}_decoder.decoder.{i + 1}.running_mean"], P[f"{mod}_decoder.decoder.{i + 1}.running_var"],
                             P[f"{mod}_decoder.decoder.{i + 1}.weight"], P[f"{mod}_decoder.decoder.{i + 1}.bias"], False, 0.0, 1e-5)
            y = F.relu(y)
        psm.append(F.conv2d(y, P[f"{mod}_cls_head.weight"], P[f"{mod}_cls_head.bias"]))
        rm.append(F.conv2d(y, P[f"{mod}_reg_head.weight"], P[f"{mod}_reg_head.bias"]))
    return torch.cat(psm, 0), torch.cat(rm, 0)


# ----------------------------------------------------------------------------------------------
# model glue: restatement of base_camera_lidar_intermediate.py, loops as in the reference
# ----------------------------------------------------------------------------------------------
def unpad_mode_encoding(mode: Tensor, record_len: Tensor) -> Tensor:
    """base_camera_lidar_intermediate.py:68-73."""
    return torch.cat([mode[i, :int(record_len[i])] for i in range(mode.shape[0])], dim=0)


def combine_features(camera_feature, lidar_feature, mode: Tensor, record_len: Tensor) -> Tensor:
    """base_camera_lidar_intermediate.py:81-99."""
    if mode.dim() == 2:
        mode = unpad_mode_encoding(mode, record_len)
    out, cc, lc = [], 0, 0
    for i in range(len(mode)):
        if mode[i] == 0:
            out.append(camera_feature[cc]); cc += 1
        elif mode[i] == 1:
            out.append(lidar_feature[lc]); lc += 1
        else:
            raise ValueError("Mode but be either 1 or 0")
    return torch.stack(out, dim=0)


def extract_lidar_input(processed_lidar: Dict[str, Tensor], mode_unpack: Tensor) -> Dict[str, Tensor]:
    """base_camera_lidar_intermediate.py:31-66 (on a copy of voxel_coords: the reference renumbers in place)."""
    coords = processed_lidar['voxel_coords'].clone()
    feats, cs, nums, count = [], [], [], 0
    for i in range(len(mode_unpack)):
        if mode_unpack[i] != 1:
            continue
        m = processed_lidar['voxel_coords'][:, 0] == i
        c = coords[m, :].clone()
        c[:, 0] = count
        cs.append(c)
        feats.append(processed_lidar['voxel_features'][m, :])
        nums.append(processed_lidar['voxel_num_points'][m])
        count += 1
    return {'voxel_features': torch.cat(feats, 0), 'voxel_coords': torch.cat(cs, 0), 'voxel_num_points': torch.cat(nums, 0)}


def detector_forward_features(camera_features, lidar_features, mode, record_len, pairwise_t_matrix, P, PD, cfg):
    """bevformer_point_pillar_hetero.py:113-134 behind the encoders: combine -> regroup -> HeteroFusion -> HeteroDecoder."""
    mode = mode.to(torch.int)
    x = combine_features(camera_features, lidar_features, unpad_mode_encoding(mode, record_len), record_len)
    x, mask = regroup(x, record_len, mode.shape[1])
    fused = hetero_fusion(x, pairwise_t_matrix, mode, record_len, mask, P, cfg)
    return hetero_decoder(fused, mode[:, 0], PD)
