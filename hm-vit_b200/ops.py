"""Thin Python wrappers over the C-ABI entry points (one per kernel family).

torch is used for device memory and the current stream only; every arithmetic step of the fusion
path happens inside libhmvit_b200.so.  All tensors must be CUDA, contiguous.
"""
import ctypes as C

import torch

from . import _lib

C_DIM = 256


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (the B200 path has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


def rowgemm(variant, *, B, L, N, n_out, mode, record_len, a, w0, w1, bias, out, ego_only=False,
            ln_gamma=None, ln_beta=None, ln_eps=1e-5, resid=None, ln_stats=None):
    args = _lib.RowGemmArgs()
    args.B, args.L, args.N, args.n_out = B, L, N, n_out
    args.mode = mode.data_ptr()
    args.record_len = record_len.data_ptr()
    args.ego_only = 1 if ego_only else 0
    args.a = a.data_ptr()
    args.w[0], args.w[1] = w0.data_ptr(), w1.data_ptr()
    args.bias = bias.data_ptr()
    args.ln_gamma = ln_gamma.data_ptr() if ln_gamma is not None else None
    args.ln_beta = ln_beta.data_ptr() if ln_beta is not None else None
    args.ln_eps = ln_eps
    args.resid = resid.data_ptr() if resid is not None else None
    args.out = out.data_ptr()
    args.ln_stats = ln_stats.data_ptr() if ln_stats is not None else None
    _lib.check(_lib.load().hmvit_rowgemm(variant, C.byref(args), _stream()))
    return out


def _as_half(w):
    """The chain kernel takes fp16 feed-forward weights (same significand width as tf32).  fp16 tensors pass through
    (the packed dicts of fusion.py / training.py carry them as w1h_* / w2h_*); fp32 ones are converted per call."""
    return w if w.dtype == torch.float16 else w.detach().half().contiguous()


def out_ffn_chain(*, B, L, N, mode, record_len, o, resid, out, wa0, wa1, ba, w1_0, w1_1, b1, w2_0, w2_1, b2,
                  ln_gamma=None, ln_beta=None, ego_only=False, ln_eps=1e-5, stats_out=None):
    w1_0, w1_1, w2_0, w2_1 = _as_half(w1_0), _as_half(w1_1), _as_half(w2_0), _as_half(w2_1)
    args = _lib.ChainArgs()
    args.B, args.L, args.N = B, L, N
    args.mode, args.record_len = mode.data_ptr(), record_len.data_ptr()
    args.ego_only = 1 if ego_only else 0
    args.o, args.resid, args.out = o.data_ptr(), resid.data_ptr(), out.data_ptr()
    args.wa[0], args.wa[1] = wa0.data_ptr(), wa1.data_ptr()
    args.ba, args.ln_eps = ba.data_ptr(), ln_eps
    args.ln_gamma = ln_gamma.data_ptr() if ln_gamma is not None else None
    args.ln_beta = ln_beta.data_ptr() if ln_beta is not None else None
    args.stats_out = stats_out.data_ptr() if stats_out is not None else None
    args.w1[0], args.w1[1], args.b1 = w1_0.data_ptr(), w1_1.data_ptr(), b1.data_ptr()
    args.w2[0], args.w2[1], args.b2 = w2_0.data_ptr(), w2_1.data_ptr(), b2.data_ptr()
    _lib.check(_lib.load().hmvit_out_ffn_chain(C.byref(args), _stream()))
    return out


def ffn_head(*, B, L, N, mode, record_len, x, w1_0, w1_1, b1, w2_0, w2_1, b2, out):
    """Typed feed-forward head on slot 0 of every scene (HeteroFusion.mlp_head,
    bevformer_point_pillar_hetero.py:46-48): out[B,256,N] = W2 gelu(W1 x + b1) + b2."""
    w1_0, w1_1, w2_0, w2_1 = _as_half(w1_0), _as_half(w1_1), _as_half(w2_0), _as_half(w2_1)
    args = _lib.HeadArgs()
    args.B, args.L, args.N = B, L, N
    args.mode, args.record_len = mode.data_ptr(), record_len.data_ptr()
    args.x, args.out = x.data_ptr(), out.data_ptr()
    args.w1[0], args.w1[1], args.b1 = w1_0.data_ptr(), w1_1.data_ptr(), b1.data_ptr()
    args.w2[0], args.w2[1], args.b2 = w2_0.data_ptr(), w2_1.data_ptr(), b2.data_ptr()
    _lib.check(_lib.load().hmvit_ffn_head(C.byref(args), _stream()))
    return out


_ATTN_WS = {}


def _attn_workspace(impl, B, L, H, W, device):
    """Scratch of hmvit_group_attn (key records of the fused kernel / compacted tiles of the split form), cached per
    implementation, shape and device."""
    key = (impl, B, L, H, W, str(device))
    ws = _ATTN_WS.get(key)
    if ws is None:
        _ATTN_WS.clear()
        nbytes = int(_lib.load().hmvit_group_attn_workspace_bytes(impl, B, L, H, W))
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _ATTN_WS[key] = ws
    return ws


_IMPL = {None: _lib.ATTN_FUSED, "fused": _lib.ATTN_FUSED, "split": _lib.ATTN_SPLIT, "single": _lib.ATTN_SINGLE}


def group_attn(*, B, L, H, W, kind, mode, record_len, cav_mask, T, cell, q, k, v, bk, bv, bias_table, out,
               ego_only=False, key_mask=None, lse=None, impl=None, workspace=None, records_valid=False):
    """impl: None / "fused" (default: key-record pass + persistent tcgen05 kernel, csrc/attn_fused.cuh), "split"
    (compaction pass + mma.sync dense pass) or "single" (one mma.sync kernel).  Shapes the fused kernel does not
    handle (more than 8 agents per scene) run the single kernel.  workspace: caller-owned scratch (uint8, at least
    attn_workspace_bytes); records_valid=True reuses the key records an earlier fused call left in it (same geometry
    and kind) instead of recomputing them."""
    args = _lib.AttnArgs()
    args.impl = _IMPL[impl]
    if args.impl != _lib.ATTN_SINGLE:
        ws = workspace if workspace is not None else _attn_workspace(args.impl, B, L, H, W, q.device)
        args.workspace, args.workspace_bytes = ws.data_ptr(), ws.numel()
        args.records_valid = 1 if (records_valid and workspace is not None) else 0
    args.B, args.L, args.H, args.W = B, L, H, W
    args.kind = kind
    args.ego_only = 1 if ego_only else 0
    args.mode, args.record_len, args.cav_mask = mode.data_ptr(), record_len.data_ptr(), cav_mask.data_ptr()
    args.T = T.data_ptr()
    args.cell = float(cell)
    args.q, args.k, args.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    args.bk, args.bv, args.bias_table = bk.data_ptr(), bv.data_ptr(), bias_table.data_ptr()
    args.key_mask = key_mask.data_ptr() if key_mask is not None else None
    args.out = out.data_ptr()
    args.lse = lse.data_ptr() if lse is not None else None
    _lib.check(_lib.load().hmvit_group_attn(C.byref(args), _stream()))
    return out


def attn_records(*, B, L, H, W, kind, mode, record_len, cav_mask, T, cell, workspace, key_mask=None):
    """The fused attention's key-record pass on its own (see hmvit_attn_records): fills `workspace` for later
    group_attn(..., workspace=workspace, records_valid=True) calls of the same geometry and kind."""
    args = _lib.AttnArgs()
    args.B, args.L, args.H, args.W, args.kind = B, L, H, W, kind
    args.mode, args.record_len, args.cav_mask = mode.data_ptr(), record_len.data_ptr(), cav_mask.data_ptr()
    args.T, args.cell = T.data_ptr(), float(cell)
    args.key_mask = key_mask.data_ptr() if key_mask is not None else None
    args.workspace, args.workspace_bytes = workspace.data_ptr(), workspace.numel()
    _lib.check(_lib.load().hmvit_attn_records(C.byref(args), _stream()))
    return workspace


# ---- backward pass -------------------------------------------------------------------------------
def bwd_row_stats(x, stats, *, B, L, N, record_len, ego_only=False, eps=1e-5):
    """stats[a*N + tok] = (mean, rstd) of the 256 channels of x (cm fp32)."""
    _lib.check(_lib.load().hmvit_bwd_row_stats(x.data_ptr(), stats.data_ptr(), B, L, N, record_len.data_ptr(),
                                               1 if ego_only else 0, eps, _stream()))
    return stats


def bwd_layernorm(dz, x, stats, dres, dx, *, B, L, N, record_len, ego_only=False):
    """dx = dres + rstd (dz - mean(dz) - z mean(dz z)); cm fp32; dx may alias dres."""
    _lib.check(_lib.load().hmvit_bwd_layernorm(dz.data_ptr(), x.data_ptr(), stats.data_ptr(), dres.data_ptr(), dx.data_ptr(),
                                               B, L, N, record_len.data_ptr(), 1 if ego_only else 0, _stream()))
    return dx


def bwd_gelu(hp, dh):
    """in place: hp <- gelu(hp), dh <- dh * gelu'(hp)."""
    _lib.check(_lib.load().hmvit_bwd_gelu(hp.data_ptr(), dh.data_ptr(), hp.numel(), _stream()))


def bwd_cast_bf16(src, dst):
    _lib.check(_lib.load().hmvit_bwd_cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), _stream()))
    return dst


def bwd_cast_colsum(src, dst, db, *, B, L, N, mode):
    """dst = bf16(src) for the five gradient planes src (5, B*L*N, 256) fp32 and, in the same pass, their typed column sums:
    db[type, p*256 + c] += sum over the rows of agents of that type (db (2, >= 1280) fp32).  Rows of padded slots must be 0."""
    _lib.check(_lib.load().hmvit_bwd_cast_colsum(src.data_ptr(), dst.data_ptr(), db.data_ptr(), db.stride(0), B, L, N,
                                                 mode.data_ptr(), _stream()))
    return dst


def bwd_colsum(y, db, *, B, L, N, mode, record_len, ego_only=False):
    """db[type] += sum over tokens of y; y cm fp32 (B*L, 256, N) or bf16 rows (B*L*N, 256); db (2, >=256) fp32 view whose
    first 256 columns are accumulated (row stride = db.stride(0))."""
    _lib.check(_lib.load().hmvit_bwd_colsum(y.data_ptr(), 1 if y.dtype == torch.bfloat16 else 0, db.data_ptr(), db.stride(0),
                                            B, L, N, mode.data_ptr(), record_len.data_ptr(), 1 if ego_only else 0, _stream()))
    return db


def bwd_wgrad(a, b, dw, *, B, L, N, mode, record_len, ego_only=False, b_stats=None, row0=0):
    """dw[type, row0 + m, n] += sum_tok a(tok, m) b(tok, n); a / b: cm fp32 or bf16 rows; dw (2, rows, 256) fp32."""
    args = _lib.WgradArgs()
    args.B, args.L, args.N = B, L, N
    args.mode, args.record_len = mode.data_ptr(), record_len.data_ptr()
    args.ego_only = 1 if ego_only else 0
    args.a, args.a_rows_bf16 = a.data_ptr(), 1 if a.dtype == torch.bfloat16 else 0
    args.b, args.b_rows_bf16 = b.data_ptr(), 1 if b.dtype == torch.bfloat16 else 0
    args.b_stats = b_stats.data_ptr() if b_stats is not None else None
    args.dw, args.dw_rows, args.dw_row0 = dw.data_ptr(), dw.shape[1], row0
    _lib.check(_lib.load().hmvit_bwd_wgrad(C.byref(args), _stream()))
    return dw


def bwd_dgrad_cat(dcat, w0, w1, out, *, B, L, N, mode, record_len):
    """out[a] (256, N) = sum_p dcat[p, rows of a] @ w[type][p*256:(p+1)*256].T, transposed to cm: the input gradient of the fused
    Q | K' | V' projection as ONE K = 1280 GEMM.  dcat bf16 (5, B*L*N, 256); w0 / w1 bf16 (1280, 256); out cm fp32."""
    _lib.check(_lib.load().hmvit_bwd_dgrad_cat(dcat.data_ptr(), w0.data_ptr(), w1.data_ptr(), out.data_ptr(), B, L, N,
                                               mode.data_ptr(), record_len.data_ptr(), _stream()))
    return out


def group_attn_bwd(*, B, L, H, W, kind, mode, record_len, cav_mask, T, cell, q, k, v, bk, bv, bias_table, o, d_o, lse,
                   dq, dk, dv, dbk, dbv, dbias_table, ego_only=False):
    args = _lib.AttnBwdArgs()
    args.B, args.L, args.H, args.W = B, L, H, W
    args.kind, args.ego_only = kind, 1 if ego_only else 0
    args.mode, args.record_len, args.cav_mask = mode.data_ptr(), record_len.data_ptr(), cav_mask.data_ptr()
    args.T, args.cell = T.data_ptr(), float(cell)
    args.q, args.k, args.v = q.data_ptr(), k.data_ptr(), v.data_ptr()
    args.bk, args.bv, args.bias_table = bk.data_ptr(), bv.data_ptr(), bias_table.data_ptr()
    args.o, args.d_o, args.lse = o.data_ptr(), d_o.data_ptr(), lse.data_ptr()
    args.dq, args.dk, args.dv = dq.data_ptr(), dk.data_ptr(), dv.data_ptr()
    args.dbk, args.dbv, args.dbias_table = dbk.data_ptr(), dbv.data_ptr(), dbias_table.data_ptr()
    _lib.check(_lib.load().hmvit_group_attn_bwd(C.byref(args), _stream()))


def dropout(a, out, *, B, L, N, record_len, seed, stream_id, p, resid=None, ego_only=False):
    """out = resid + keep * a / (1 - p) on cm fp32 (B*L, 256, N) tensors; a=None exports the scaled mask.  The mask is a
    pure function of (seed, stream_id, element index): the backward regenerates it with the same arguments."""
    _lib.check(_lib.load().hmvit_dropout(a.data_ptr() if a is not None else None, resid.data_ptr() if resid is not None else None,
                                         out.data_ptr(), B, L, N, record_len.data_ptr(), 1 if ego_only else 0,
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF, float(p), _stream()))
    return out


def warp_bilinear(x: torch.Tensor, T: torch.Tensor, cell: float) -> torch.Tensor:
    """x (n, C, H, W) fp32, T (n, 4, 4) fp32 source->target.  Returns the warped maps (n, C, H, W)."""
    _chk(x, torch.float32, "x"), _chk(T, torch.float32, "T")
    n, Cc, H, W = x.shape
    if T.shape != (n, 4, 4):
        raise ValueError(f"T: expected shape {(n, 4, 4)}, got {tuple(T.shape)}")
    out = torch.empty_like(x)
    _lib.check(_lib.load().hmvit_warp_bilinear(x.data_ptr(), T.data_ptr(), out.data_ptr(), n, Cc, H, W, float(cell), _stream()))
    return out


def roi_cav_mask(T: torch.Tensor, cav_mask: torch.Tensor, H: int, W: int, cell: float) -> torch.Tensor:
    """T (B, L, 4, 4) fp32, cav_mask (B, L) int32 -> float32 (B, H, W, 1, L)."""
    _chk(T, torch.float32, "T"), _chk(cav_mask, torch.int32, "cav_mask")
    B, L = T.shape[:2]
    out = torch.empty(B, H, W, 1, L, dtype=torch.float32, device=T.device)
    _lib.check(_lib.load().hmvit_roi_cav_mask(T.data_ptr(), cav_mask.data_ptr(), out.data_ptr(), B, L, H, W, float(cell), _stream()))
    return out


def attn_workspace_bytes(B, L, H, W, impl=None) -> int:
    return int(_lib.load().hmvit_group_attn_workspace_bytes(_IMPL[impl], B, L, H, W))


def fusion_workspace_bytes(B, L, H, W, unfused=False, attn_impl=None) -> int:
    return int(_lib.load().hmvit_fusion_workspace_bytes(B, L, H, W, 1 if unfused else 0, _IMPL[attn_impl]))


def fusion_launch_count(num_iters, head, skip_dead=True, attn_impl=None) -> int:
    """Kernel launches of one hmvit_fusion_forward.  With skip_dead (the module default) the head runs inside the last
    stage's chain launch instead of a launch of its own."""
    return int(_lib.load().hmvit_fusion_launch_count(num_iters, (2 if skip_dead else 1) if head else 0, _IMPL[attn_impl]))


def decoder_workspace_bytes(B, H, W):
    return int(_lib.load().hmvit_decoder_workspace_bytes(B, H, W))


def decoder_forward(*, x, ego_mode, conv_w, conv_b, head_w, head_b, anchor_number, psm, rm, workspace):
    """HeteroDecoder.forward(use_upsample=False) on the ego's fused feature x (B, 256, H, W) fp32: BatchNorm-folded fp16
    convolution weights [num_convs][2][9][256][256], fp32 biases / head weights (see include/hmvit_b200.h)."""
    args = _lib.DecoderArgs()
    B, _, H, W = x.shape
    args.B, args.H, args.W = B, H, W
    args.num_convs = conv_w.shape[0]
    args.anchor_number = anchor_number
    args.ego_mode, args.x = ego_mode.data_ptr(), x.data_ptr()
    args.conv_w, args.conv_b = conv_w.data_ptr(), conv_b.data_ptr()
    args.head_w, args.head_b = head_w.data_ptr(), head_b.data_ptr()
    args.psm, args.rm = psm.data_ptr(), rm.data_ptr()
    args.workspace, args.workspace_bytes = workspace.data_ptr(), workspace.numel()
    _lib.check(_lib.load().hmvit_decoder_forward(C.byref(args), _stream()))
    return psm, rm


def postprocess(*, psm, rm, anchor_box, transformation_matrix, order_hwl, score_threshold, nms_thresh, gt_range):
    """VoxelPostprocessor.post_process for one cav, batch 1 (include/hmvit_b200.h): returns device tensors
    (boxes [1000][8][3], scores [1000], meta [3] int32 = (boxes kept, status, candidates before the NMS))."""
    A, H, W = psm.shape[1], psm.shape[2], psm.shape[3]
    dev = psm.device
    lib = _lib.load()
    nbytes = int(lib.hmvit_postprocess_workspace_bytes(H, W, A))
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    off = (-ws.data_ptr()) % 256
    boxes = torch.empty(1000, 8, 3, dtype=torch.float32, device=dev)
    scores = torch.empty(1000, dtype=torch.float32, device=dev)
    meta = torch.zeros(3, dtype=torch.int32, device=dev)
    args = _lib.PostArgs()
    args.H, args.W, args.A = H, W, A
    args.psm, args.rm, args.anchor_box = psm.data_ptr(), rm.data_ptr(), anchor_box.data_ptr()
    args.transformation_matrix = transformation_matrix.data_ptr() if transformation_matrix is not None else None
    args.order_hwl = 1 if order_hwl else 0
    args.score_threshold, args.nms_thresh = float(score_threshold), float(nms_thresh)
    args.range[0], args.range[1], args.range[2], args.range[3] = gt_range[0], gt_range[1], gt_range[3], gt_range[4]
    args.out_boxes, args.out_scores = boxes.data_ptr(), scores.data_ptr()
    args.out_count, args.status = meta[0:1].data_ptr(), meta[1:2].data_ptr()
    args.workspace, args.workspace_bytes = ws.data_ptr() + off, nbytes
    _lib.check(lib.hmvit_postprocess(C.byref(args), _stream()))
    # the candidate count lives in the last 256 bytes of the workspace (csrc/api.cu)
    meta[2:3].copy_(ws[off + nbytes - 256: off + nbytes - 252].view(torch.int32))
    return boxes, scores, meta


def pillar_scatter(*, voxel_features, voxel_coords, voxel_num_points, w, b, voxel_size, offset, nx, ny, n_agents,
                   channels_last=False):
    """PillarVFE (one PFN layer, BatchNorm folded into w / b) + PointPillarScatter in one kernel (include/hmvit_b200.h):
    voxels -> canvas (n_agents, 64, ny, nx) fp32; with channels_last the same logical tensor in torch's channels_last
    memory format."""
    dev = voxel_features.device
    M, P = int(voxel_features.shape[0]), int(voxel_features.shape[1])
    if channels_last:
        canvas = torch.empty(n_agents, ny, nx, 64, dtype=torch.float32, device=dev).permute(0, 3, 1, 2)
    else:
        canvas = torch.empty(n_agents, 64, ny, nx, dtype=torch.float32, device=dev)
    args = _lib.PillarArgs()
    args.M, args.P = M, P
    args.voxel_features, args.voxel_coords, args.voxel_num_points = (voxel_features.data_ptr(), voxel_coords.data_ptr(),
                                                                     voxel_num_points.data_ptr())
    args.w, args.b = w.data_ptr(), b.data_ptr()
    for k in range(3):
        args.voxel_size[k], args.offset[k] = float(voxel_size[k]), float(offset[k])
    args.nx, args.ny, args.n_agents, args.canvas = nx, ny, n_agents, canvas.data_ptr()
    args.channels_last = 1 if channels_last else 0
    _lib.check(_lib.load().hmvit_pillar_scatter(C.byref(args), _stream()))
    return canvas
