"""Drop-in host side of the HM-ViT fusion path: the reference's nn.Module surface over the sm_100a
C-ABI (libhmvit_b200.so).

Mirrors, name for name (constructor config, forward signature, state_dict keys and shapes):
  HeteroFusion        /root/reference/opencood/models/bevformer_point_pillar_hetero.py:22-49
  HeteroFusionBlock   /root/reference/opencood/models/sub_modules/hetero_fusion.py:279-474
  HeteroAttention     /root/reference/opencood/models/sub_modules/hetero_fusion.py:16-277
  HeteroLayerNorm / HeteroFeedForward / HeteroPreNormResidual   .../models/base_transformer.py:121-192
  SpatialTransformation   .../sub_modules/spatial_transformation.py:10-44
  get_roi_and_cav_mask    .../sub_modules/torch_transformation_utils.py:11-49
  regroup                 .../sub_modules/fuse_utils.py:8-61
so that `net_epoch*.pth` checkpoints load unchanged (prefix `fusion_net.` in the detector).

PyTorch here is plumbing: parameter storage, device memory, streams.  All arithmetic of the forward
runs in hand-written CUDA kernels; a missing extension raises (no eager fallback).  With grad enabled the
module runs the training path (hmvit_b200/training.py: activations saved per stage, hand-written backward).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch
from torch import nn

from . import _lib, ops, training

_NUM_TYPES = 2
_LOG2E = 1.4426950408889634


# ----------------------------------------------------------------------------------------------
# parameter containers (same registration order / names as the reference)
# ----------------------------------------------------------------------------------------------
class HeteroLayerNorm(nn.Module):
    """Per-modality LayerNorm parameters (base_transformer.py:171-177).  Evaluated inside the fused
    row-GEMM prologue, not as a separate op."""

    def __init__(self, dim, num_types=_NUM_TYPES):
        super().__init__()
        self.num_types = num_types
        self.net = nn.ModuleList([nn.LayerNorm(dim) for _ in range(num_types)])


class HeteroFeedForward(nn.Module):
    """Per-modality Linear-GELU-Dropout-Linear-Dropout parameters (base_transformer.py:180-192)."""

    def __init__(self, dim, hidden_dim, dropout=0., num_types=_NUM_TYPES, out_dim=None):
        super().__init__()
        out_dim = dim if out_dim is None else out_dim
        self.num_types = num_types
        self.net = nn.ModuleList([
            nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                          nn.Linear(hidden_dim, out_dim), nn.Dropout(dropout))
            for _ in range(num_types)])


class HeteroPreNormResidual(nn.Module):
    def __init__(self, dim, fn, num_types=_NUM_TYPES):
        super().__init__()
        self.norm = HeteroLayerNorm(dim, num_types=num_types)
        self.fn = fn


class SpatialTransformation(nn.Module):
    """spatial_transformation.py:10-44 -- bilinear warp of every agent map with its own 4x4 pose."""

    def __init__(self, args):
        super().__init__()
        self.discrete_ratio = args['voxel_size'][0]
        self.downsample_rate = args['downsample_rate']

    def forward(self, x, spatial_correction_matrix):
        B, L, Cc, H, W = x.shape
        y = ops.warp_bilinear(x.reshape(B * L, Cc, H, W).float().contiguous(),
                              spatial_correction_matrix.reshape(B * L, 4, 4).float().contiguous(),
                              float(self.discrete_ratio) * float(self.downsample_rate))
        return y.reshape(B, L, Cc, H, W)


def get_roi_and_cav_mask(shape, cav_mask, spatial_correction_matrix, discrete_ratio, downsample_rate):
    """torch_transformation_utils.py:11-49 -> float32 (B, H, W, 1, L)."""
    B, L, H, W, _ = shape
    return ops.roi_cav_mask(spatial_correction_matrix.float().contiguous(), cav_mask.to(torch.int32).contiguous(),
                            H, W, float(discrete_ratio) * float(downsample_rate))


def regroup(dense_feature, record_len, max_len):
    """fuse_utils.py:8-61: (sum L_b, C, H, W) -> zero padded (B, L, C, H, W) + int64 (B, L) mask.
    Like the reference this reads record_len on the host."""
    lens = [int(v) for v in record_len.tolist()]
    _, Cc, H, W = dense_feature.shape
    out = dense_feature.new_zeros(len(lens), max_len, Cc, H, W)
    mask = torch.zeros(len(lens), max_len, dtype=torch.int64)
    start = 0
    for b, n in enumerate(lens):
        out[b, :n] = dense_feature[start:start + n]
        mask[b, :n] = 1
        start += n
    return out, mask.to(dense_feature.device)


# ----------------------------------------------------------------------------------------------
# weight packing (one-off per parameter version: folds scale / edge-type weights, casts)
# ----------------------------------------------------------------------------------------------
def _tf32_round(w: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest (ties away) fp32 -> tf32, kept in fp32 storage (matches cvt.rna.tf32.f32)."""
    i = w.detach().float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _stack2(fn):
    return torch.stack([fn(0), fn(1)]).contiguous()


class HeteroAttention(nn.Module):
    """Parameters of the typed multi-agent attention (hetero_fusion.py:16-109) + the unit-level forward."""

    attn_impl = None      # None: fused persistent tcgen05 attention; "split" / "single": cross-check forms (ops.group_attn)

    def __init__(self, dim, dim_head=32, dropout=0., agent_size=6, window_size=7, num_types=_NUM_TYPES):
        super().__init__()
        assert (dim % dim_head) == 0, 'dimension should be divisible by dimension per head'
        self.heads = dim // dim_head
        self.scale = dim_head ** -0.5
        self.num_types = num_types
        self.use_position_emb = True
        self.window_size = [agent_size, window_size, window_size]
        self.attend = nn.Sequential(nn.Softmax(dim=-1))
        self.k_linears, self.q_linears, self.v_linears = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.a_linears, self.norms = nn.ModuleList(), nn.ModuleList()
        for _ in range(num_types):
            self.k_linears.append(nn.Linear(dim, dim))
            self.q_linears.append(nn.Linear(dim, dim))
            self.v_linears.append(nn.Linear(dim, dim))
            self.a_linears.append(nn.Sequential(nn.Linear(dim, dim), nn.Dropout(dropout)))
        num_relations = num_types ** 2
        self.relation_att = nn.Parameter(torch.empty(num_relations, self.heads, dim_head, dim_head))
        self.relation_msg = nn.Parameter(torch.empty(num_relations, self.heads, dim_head, dim_head))
        Wh = Ww = window_size
        self.relative_position_bias_table = nn.Embedding((2 * Wh - 1) * (2 * Ww - 1), self.heads)
        r = torch.arange(Wh).view(Wh, 1).expand(Wh, Ww).reshape(-1)
        c = torch.arange(Ww).view(1, Ww).expand(Wh, Ww).reshape(-1)
        idx = (r[:, None] - r[None, :] + Wh - 1) * (2 * Ww - 1) + (c[:, None] - c[None, :] + Ww - 1)
        self.register_buffer("relative_position_index", idx)
        nn.init.xavier_uniform_(self.relation_att)
        nn.init.xavier_uniform_(self.relation_msg)
        self._dim, self._dim_head = dim, dim_head
        _check_supported(dim, dim_head, window_size)

    # folded projection weights, see DESIGN.md "exact restructurings"
    def packed(self, ln_gamma=None, ln_beta=None) -> Dict[str, torch.Tensor]:
        """Folded projection weights.  With ln_gamma / ln_beta ([2][C], per type) the affine of the
        preceding typed LayerNorm is folded in as well:  W (g*z + b) + c = (W diag g) z + (W b + c)."""
        Cd, h, d = self._dim, self.heads, self._dim_head
        # the fused attention kernel bounds its probabilities by 2^(8 + max|bias| log2 e): keep that inside fp32 range
        bmax = float(self.relative_position_bias_table.weight.detach().abs().max())
        if not bmax <= 40.0:
            raise ValueError(f"relative_position_bias_table: |bias| up to {bmax:.1f} exceeds the supported range of +-40")
        att, msg = self.relation_att.detach().float(), self.relation_msg.detach().float()
        wqkv, bqkv = [], []
        bk = torch.empty(2, 2, Cd, device=att.device)
        bv = torch.empty(2, 2, Cd, device=att.device)
        for t in range(2):
            # softmax scale and log2(e) folded into the query projection (the kernel works in the exp2 domain)
            wq = self.q_linears[t].weight.detach().float() * (self.scale * _LOG2E)
            bq = self.q_linears[t].bias.detach().float() * (self.scale * _LOG2E)
            wk = self.k_linears[t].weight.detach().float().view(h, d, Cd)
            bkt = self.k_linears[t].bias.detach().float().view(h, d)
            wv = self.v_linears[t].weight.detach().float().view(h, d, Cd)
            bvt = self.v_linears[t].bias.detach().float().view(h, d)
            parts = [wq]
            for te in range(2):
                e = te * 2 + t
                parts.append(torch.einsum("hpq,hqc->hpc", att[e], wk).reshape(Cd, Cd))
                bk[te, t] = torch.einsum("hpq,hq->hp", att[e], bkt).reshape(Cd)
            for te in range(2):
                e = te * 2 + t
                parts.append(torch.einsum("hpq,hpc->hqc", msg[e], wv).reshape(Cd, Cd))
                bv[te, t] = torch.einsum("hpq,hp->hq", msg[e], bvt).reshape(Cd)
            wcat = torch.cat(parts, 0)                                   # [5C, C] fp32
            bcat = torch.cat([bq, bq.new_zeros(4 * Cd)])
            if ln_gamma is not None:
                bcat = bcat + wcat @ ln_beta[t].float()
                wcat = wcat * ln_gamma[t].float()[None, :]
            wqkv.append(wcat.to(torch.bfloat16).contiguous())
            bqkv.append(bcat)
        return {
            "wqkv0": wqkv[0], "wqkv1": wqkv[1], "bqkv": torch.stack(bqkv).contiguous(),
            "bk": bk.contiguous(), "bv": bv.contiguous(),
            "wa0": self.a_linears[0][0].weight.detach().to(torch.bfloat16).contiguous(),
            "wa1": self.a_linears[1][0].weight.detach().to(torch.bfloat16).contiguous(),
            "ba": _stack2(lambda t: self.a_linears[t][0].bias.detach().float()),
            "bias_table": self.relative_position_bias_table.weight.detach().float().contiguous(),
        }

    def forward(self, x, mode, mask=None, exclude_self=False):
        """Unit-level surface (hetero_fusion.py:187-277): x (b, l, X, Y, w, w, c) already normalised,
        warped into the ego frame (ego = slot 0) and window-partitioned; mask (b, X, Y, w, w, 1, l).
        Returns (b, 1, X, Y, w, w, c)."""
        if exclude_self:
            raise NotImplementedError("exclude_self=True is never used by HM-ViT (hetero_fusion.py:454,458)")
        _require_inference(self)
        b, l, X, Y, w1, w2, c = x.shape
        H, W = X * w1, Y * w2
        dev = x.device
        xim = x.permute(0, 1, 6, 2, 4, 3, 5).reshape(b, l, c, H, W).float().contiguous()
        N = H * W
        mode_i = mode.to(torch.int32).contiguous()
        rl = torch.full((b,), l, dtype=torch.int32, device=dev)
        cav = torch.ones(b, l, dtype=torch.int32, device=dev)
        T = torch.eye(4, device=dev).repeat(b, l, l, 1, 1).contiguous()
        km = None
        if mask is not None:
            km = (mask != 0).permute(0, 6, 5, 1, 3, 2, 4).reshape(b, l, N).to(torch.uint8).contiguous()
        pk = self.packed()
        rows = b * l * N
        qkv = torch.empty(5, rows, c, dtype=torch.bfloat16, device=dev)
        ops.rowgemm(_lib.GEMM_QKV_NOLN, B=b, L=l, N=N, n_out=5 * c, mode=mode_i, record_len=rl, a=xim,
                    w0=pk["wqkv0"], w1=pk["wqkv1"], bias=pk["bqkv"], out=qkv, ego_only=True)
        att = torch.zeros(rows, c, dtype=torch.bfloat16, device=dev)
        ops.group_attn(B=b, L=l, H=H, W=W, kind=0, mode=mode_i, record_len=rl, cav_mask=cav, T=T, cell=1.0,
                       q=qkv[0], k=qkv[1:3], v=qkv[3:5], bk=pk["bk"], bv=pk["bv"], bias_table=pk["bias_table"],
                       out=att, ego_only=True, key_mask=km, impl=self.attn_impl)
        zero = torch.zeros(b, l, c, N, dtype=torch.float32, device=dev)
        out = torch.zeros(b, l, c, N, dtype=torch.float32, device=dev)
        ops.rowgemm(_lib.GEMM_OUT, B=b, L=l, N=N, n_out=c, mode=mode_i, record_len=rl, a=att,
                    w0=pk["wa0"], w1=pk["wa1"], bias=pk["ba"], out=out, resid=zero, ego_only=True)
        y = out[:, :1].reshape(b, 1, c, X, w1, Y, w2).permute(0, 1, 3, 5, 4, 6, 2).contiguous()
        return y


def _check_supported(dim, dim_head, window_size):
    if dim != 256 or dim_head != 32 or window_size != 8:
        raise ValueError(
            f"hmvit_b200 kernels are specialised for input_dim=256, dim_head=32, window_size=8 "
            f"(got {dim}, {dim_head}, {window_size})")


def _require_inference(module: nn.Module):
    if torch.is_grad_enabled() and any(p.requires_grad for p in module.parameters()):
        raise NotImplementedError(
            "the unit-level HeteroAttention.forward surface is inference only: call under torch.no_grad(); "
            "gradients are implemented for HeteroFusionBlock / HeteroFusion (hmvit_b200/training.py)")


def _wants_grad(module: nn.Module, x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in module.parameters()))


def _dropout_args(block):
    """(p, seed) of this forward.  The reference applies Dropout(p) after a_linears and twice in the FFN in train mode
    (hetero_fusion.py:66, base_transformer.py:186-190); here the three sites draw from a counter-based Philox stream
    (csrc/dropout.cuh) keyed by a per-forward seed: `block.dropout_seed` when set (reproducible runs / mask export in the
    tests), else a fresh draw from torch's default generator (so torch.manual_seed governs it).  Same distribution as
    the reference, not the same bits as torch's own RNG."""
    if not (block.training and block.drop_out > 0):
        return 0.0, 0
    seed = block.dropout_seed
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    block.last_dropout_seed = seed
    return float(block.drop_out), seed


def _check_inputs(block, x, pairwise_t_matrix, mode, record_len, mask):
    """Shape / device validation shared by the inference and the training entry of the module.  (record_len values are
    clamped to L inside the kernels, so a malformed value cannot index past a scene's slots without a host sync here.)"""
    if x.dim() != 5:
        raise ValueError(f"x: expected (B, L, C, H, W), got {tuple(x.shape)}")
    if not x.is_cuda:
        raise ValueError("hmvit_b200 runs on CUDA tensors only (no CPU fallback)")
    B, L, Cc, H, W = x.shape
    if Cc != 256:
        raise ValueError(f"x: channel dim must be 256, got {Cc}")
    if H % block.window_size or W % block.window_size:
        raise ValueError(f"H={H}, W={W} must be divisible by the window size {block.window_size}")
    if tuple(pairwise_t_matrix.shape) != (B, L, L, 4, 4):
        raise ValueError(f"pairwise_t_matrix: expected {(B, L, L, 4, 4)}, got {tuple(pairwise_t_matrix.shape)}")
    if tuple(mode.shape) != (B, L) or tuple(mask.shape) != (B, L) or tuple(record_len.shape) != (B,):
        raise ValueError("mode / mask must be (B, L) and record_len (B,)")


class HeteroFusionBlock(nn.Module):
    """hetero_fusion.py:279-474 -- window stage then grid stage, each LN -> warp -> per-ego attention
    -> residual -> pre-norm FFN residual.  Only architect_mode == 'sequential' (the shipped yaml)."""

    attn_impl = None      # attention implementation inside hmvit_fusion_forward (see HeteroAttention.attn_impl)
    unfused_chain = False  # True: OUT / FFN1 / FFN2 as three row-GEMMs instead of the fused chain kernel (cross-check)
    dropout_seed = None    # int: fixed Philox seed of the train-mode Dropout (None: drawn per forward)
    last_dropout_seed = None

    def __init__(self, config):
        super().__init__()
        input_dim, mlp_dim = config['input_dim'], config['mlp_dim']
        agent_size, window_size = config['agent_size'], config['window_size']
        drop_out, dim_head = config['drop_out'], config['dim_head']
        self.architect_mode = config['architect_mode']
        self.drop_out = float(drop_out)
        if mlp_dim != input_dim:
            raise ValueError("hmvit_b200 kernels require mlp_dim == input_dim == 256")
        self.spatial_transform = SpatialTransformation(config['spatial_transform'])
        self.downsample_rate = config['spatial_transform']['downsample_rate']
        self.discrete_ratio = config['spatial_transform']['voxel_size'][0]
        self.window_size = window_size
        self.window_norm = HeteroLayerNorm(input_dim)
        self.window_attention = HeteroAttention(input_dim, dim_head, drop_out, agent_size, window_size)
        self.window_ffd = HeteroPreNormResidual(input_dim, HeteroFeedForward(input_dim, mlp_dim, drop_out))
        self.grid_norm = HeteroLayerNorm(input_dim)
        self.grid_attention = HeteroAttention(input_dim, dim_head, drop_out, agent_size, window_size)
        self.grid_ffd = HeteroPreNormResidual(input_dim, HeteroFeedForward(input_dim, mlp_dim, drop_out))
        # present in the reference state_dict but never used in forward (hetero_fusion.py:326)
        self.aggregate_fc = HeteroFeedForward(mlp_dim * 3, mlp_dim, drop_out, out_dim=mlp_dim)
        self._pack_cache = None

    # ---- packed weights, cached on parameter versions -----------------------------------------
    def _stage_pack(self, kind: str) -> Dict[str, torch.Tensor]:
        att = getattr(self, f"{kind}_attention")
        norm = getattr(self, f"{kind}_norm")
        ffd = getattr(self, f"{kind}_ffd")
        # both typed LayerNorm affines are folded into the GEMM that consumes them (exact algebra), so the
        # kernels only normalise: z = (x - mean) * rstd
        g1 = _stack2(lambda t: norm.net[t].weight.detach().float())
        b1n = _stack2(lambda t: norm.net[t].bias.detach().float())
        g2 = _stack2(lambda t: ffd.norm.net[t].weight.detach().float())
        b2n = _stack2(lambda t: ffd.norm.net[t].bias.detach().float())
        pk = att.packed(g1, b1n)
        b1 = []
        for t in range(2):
            w1 = ffd.fn.net[t][0].weight.detach().float()
            pk[f"w1_{t}"] = _tf32_round(w1 * g2[t][None, :])
            b1.append(ffd.fn.net[t][0].bias.detach().float() + w1 @ b2n[t])
            pk[f"w2_{t}"] = _tf32_round(ffd.fn.net[t][3].weight)
            # fp16 copies for the fused chain kernel: the same 11-bit significand as tf32 at twice the MMA rate
            pk[f"w1h_{t}"] = (w1 * g2[t][None, :]).half().contiguous()
            pk[f"w2h_{t}"] = ffd.fn.net[t][3].weight.detach().half().contiguous()
        pk["b1"] = torch.stack(b1).contiguous()
        pk["b2"] = _stack2(lambda t: ffd.fn.net[t][3].bias.detach().float())
        return pk

    def packed(self):
        """Folded / cast weights, cached on (data_ptr, version) of every parameter.  In-place writes through `.data`
        (EMA swaps, `p.data.copy_`) do not bump the version: call invalidate_packed() after them.  load_state_dict,
        `.to()` / `.half()` (`_apply`) and `.train()` drop the cache themselves."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._pack_cache is None or self._pack_cache[0] != key:
            self._pack_cache = (key, {"window": self._stage_pack("window"), "grid": self._stage_pack("grid")})
        return self._pack_cache[1]

    def invalidate_packed(self):
        self._pack_cache = None

    def _load_from_state_dict(self, *args, **kwargs):
        self._pack_cache = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._pack_cache = None
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode=True):
        self._pack_cache = None
        return super().train(mode)

    def forward(self, x, pairwise_t_matrix, mode, record_len, mask):
        """x (B, L, C, H, W) -> (B, L, C, H, W).  Valid slots (l < record_len[b]) hold the block output;
        padded slots are returned unchanged (the reference leaves discarded values there)."""
        if self.architect_mode != 'sequential':
            raise ValueError(f"{self.architect_mode} not implemented")
        if _wants_grad(self, x):
            _check_inputs(self, x, pairwise_t_matrix, mode, record_len, mask)
            p, seed = _dropout_args(self)
            return training.fusion_train(ops, self, None, x, pairwise_t_matrix, mode, record_len, mask, num_iters=1,
                                         drop_p=p, seed=seed)
        xres = x.detach().float().clone().contiguous()
        _run_fusion(self, None, x, pairwise_t_matrix, mode, record_len, mask, num_iters=1, xres=xres)
        return xres


class HeteroFusion(nn.Module):
    """bevformer_point_pillar_hetero.py:22-49 -- the drop-in class: num_iters x the SAME block (shared
    weights), ego slice, typed FFN head.  forward(x, pairwise_t_matrix, mode, record_len, mask) -> (B, C, H, W)."""

    def __init__(self, config):
        super().__init__()
        self.spatial_transform = SpatialTransformation(config['spatial_transform'])
        self.downsample_rate = config['spatial_transform']['downsample_rate']
        self.discrete_ratio = config['spatial_transform']['voxel_size'][0]
        self.hetero_fusion_block = HeteroFusionBlock(config['hetero_fusion_block'])
        input_dim = config['hetero_fusion_block']['input_dim']
        self.num_iters = config['num_iters']
        self.mlp_head = HeteroFeedForward(input_dim, input_dim, 0)
        self.skip_dead_queries = True      # exact: the last grid stage only needs the ego's queries
        self._head_cache = None

    def head_packed(self):
        key = tuple((p.data_ptr(), p._version) for p in self.mlp_head.parameters())
        if self._head_cache is None or self._head_cache[0] != key:
            net = self.mlp_head.net
            pk = {f"w1_{t}": _tf32_round(net[t][0].weight) for t in range(2)}
            pk.update({f"w2_{t}": _tf32_round(net[t][3].weight) for t in range(2)})
            # fp16 copies for the fused head kernel (same significand width as tf32)
            pk.update({f"w1h_{t}": net[t][0].weight.detach().half().contiguous() for t in range(2)})
            pk.update({f"w2h_{t}": net[t][3].weight.detach().half().contiguous() for t in range(2)})
            pk["b1"] = _stack2(lambda t: net[t][0].bias.detach().float())
            pk["b2"] = _stack2(lambda t: net[t][3].bias.detach().float())
            self._head_cache = (key, pk)
        return self._head_cache[1]

    def invalidate_packed(self):
        """Drop the packed-weight caches of the head and of the block (after in-place writes through `.data`)."""
        self._head_cache = None
        self.hetero_fusion_block.invalidate_packed()

    def _load_from_state_dict(self, *args, **kwargs):
        self._head_cache = None
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._head_cache = None
        return super()._apply(fn, *args, **kwargs)

    def train(self, mode=True):
        self._head_cache = None
        return super().train(mode)

    def forward(self, x, pairwise_t_matrix, mode, record_len, mask):
        blk = self.hetero_fusion_block
        if blk.architect_mode != 'sequential':
            raise ValueError(f"{blk.architect_mode} not implemented")
        if _wants_grad(self, x):
            _check_inputs(blk, x, pairwise_t_matrix, mode, record_len, mask)
            p, seed = _dropout_args(blk)
            return training.fusion_train(ops, blk, self, x, pairwise_t_matrix, mode, record_len, mask,
                                         num_iters=self.num_iters, skip_dead=self.skip_dead_queries, drop_p=p, seed=seed)
        return _run_fusion(blk, self, x, pairwise_t_matrix, mode, record_len, mask, num_iters=self.num_iters)


# ----------------------------------------------------------------------------------------------
# the one call into the C-ABI for the whole forward
# ----------------------------------------------------------------------------------------------
_WS_CACHE: Dict[Tuple, torch.Tensor] = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (str(device),)
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS_CACHE[key] = ws
    return ws


def _fill_stage(sw: "_lib.StageWeights", pk: Dict[str, torch.Tensor]):
    sw.wqkv[0], sw.wqkv[1] = pk["wqkv0"].data_ptr(), pk["wqkv1"].data_ptr()
    sw.bqkv, sw.bk, sw.bv = pk["bqkv"].data_ptr(), pk["bk"].data_ptr(), pk["bv"].data_ptr()
    sw.wa[0], sw.wa[1] = pk["wa0"].data_ptr(), pk["wa1"].data_ptr()
    sw.ba = pk["ba"].data_ptr()
    sw.ln1_g = sw.ln1_b = sw.ln2_g = sw.ln2_b = None      # LayerNorm affines are folded into the weights
    sw.w1[0], sw.w1[1] = pk["w1_0"].data_ptr(), pk["w1_1"].data_ptr()
    sw.w2[0], sw.w2[1] = pk["w2_0"].data_ptr(), pk["w2_1"].data_ptr()
    sw.b1, sw.b2 = pk["b1"].data_ptr(), pk["b2"].data_ptr()
    sw.bias_table = pk["bias_table"].data_ptr()
    sw.w1h[0], sw.w1h[1] = pk["w1h_0"].data_ptr(), pk["w1h_1"].data_ptr()
    sw.w2h[0], sw.w2h[1] = pk["w2h_0"].data_ptr(), pk["w2h_1"].data_ptr()


def _run_fusion(block: HeteroFusionBlock, fusion, x, pairwise_t_matrix, mode, record_len, mask, num_iters, xres=None):
    _check_inputs(block, x, pairwise_t_matrix, mode, record_len, mask)
    B, L, Cc, H, W = x.shape
    dev = x.device
    x = x.detach().float().contiguous()
    T = pairwise_t_matrix.detach().to(device=dev, dtype=torch.float32).contiguous()
    mode_i = mode.detach().to(device=dev, dtype=torch.int32).contiguous()
    rl = record_len.detach().to(device=dev, dtype=torch.int32).contiguous()
    cav = mask.detach().to(device=dev, dtype=torch.int32).contiguous()

    head = fusion is not None
    if xres is None:
        xres = torch.empty_like(x)
    unfused = bool(block.unfused_chain)
    attn_impl = block.attn_impl
    ws = _workspace(ops.fusion_workspace_bytes(B, L, H, W, unfused, attn_impl), dev)
    out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=dev) if head else None

    args = _lib.FusionArgs()
    args.B, args.L, args.H, args.W = B, L, H, W
    args.num_iters, args.head = num_iters, 1 if head else 0
    args.skip_dead = 1 if (head and fusion.skip_dead_queries) else 0
    args.unfused = 1 if unfused else 0
    args.attn_impl = ops._IMPL[attn_impl]
    args.x, args.T = x.data_ptr(), T.data_ptr()
    args.mode, args.record_len, args.cav_mask = mode_i.data_ptr(), rl.data_ptr(), cav.data_ptr()
    args.cell = float(block.discrete_ratio) * float(block.downsample_rate)
    args.ln_eps = 1e-5
    pk = block.packed()
    _fill_stage(args.stage[0], pk["window"])
    _fill_stage(args.stage[1], pk["grid"])
    keep = [pk]
    if head:
        hp = fusion.head_packed()
        keep.append(hp)
        args.head_w1[0], args.head_w1[1] = hp["w1_0"].data_ptr(), hp["w1_1"].data_ptr()
        args.head_w2[0], args.head_w2[1] = hp["w2_0"].data_ptr(), hp["w2_1"].data_ptr()
        args.head_w1h[0], args.head_w1h[1] = hp["w1h_0"].data_ptr(), hp["w1h_1"].data_ptr()
        args.head_w2h[0], args.head_w2h[1] = hp["w2h_0"].data_ptr(), hp["w2h_1"].data_ptr()
        args.head_b1, args.head_b2 = hp["b1"].data_ptr(), hp["b2"].data_ptr()
        args.out = out.data_ptr()
    args.xres, args.workspace = xres.data_ptr(), ws.data_ptr()
    _lib.check(_lib.load().hmvit_fusion_forward(C.byref(args), torch.cuda.current_stream().cuda_stream))
    return out if head else xres
