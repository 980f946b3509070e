"""ctypes binding of libhmvit_b200.so (C-ABI declared in include/hmvit_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
Build the library with `python -c "import __graft_entry__ as g; g.build()"` (nvcc, sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HMVIT_LIB", os.path.join(_HERE, "libhmvit_b200.so"))   # HMVIT_LIB: tuning builds only

GEMM_QKV, GEMM_OUT, GEMM_FFN1, GEMM_FFN2, GEMM_HEAD1, GEMM_HEAD2, GEMM_QKV_NOLN = range(7)
GEMM_LN_LIN_CM, GEMM_LIN_CM, GEMM_LIN_ROWS, GEMM_ROWS_LIN_CM = range(7, 11)
ATTN_FUSED, ATTN_SPLIT, ATTN_SINGLE = range(3)     # HmvitAttnArgs.impl
ABI_VERSION = 9

EXPORTS = (
    "hmvit_abi_version", "hmvit_last_error", "hmvit_rowgemm", "hmvit_group_attn", "hmvit_warp_bilinear",
    "hmvit_roi_cav_mask", "hmvit_fusion_workspace_bytes", "hmvit_fusion_forward", "hmvit_fusion_launch_count",
    "hmvit_out_ffn_chain", "hmvit_ffn_head",
    "hmvit_bwd_row_stats", "hmvit_bwd_layernorm", "hmvit_bwd_gelu", "hmvit_bwd_cast_bf16", "hmvit_bwd_colsum", "hmvit_bwd_dgrad_cat", "hmvit_bwd_cast_colsum",
    "hmvit_bwd_wgrad", "hmvit_group_attn_bwd", "hmvit_group_attn_workspace_bytes", "hmvit_dropout", "hmvit_attn_records",
    "hmvit_decoder_workspace_bytes", "hmvit_decoder_forward", "hmvit_postprocess_workspace_bytes", "hmvit_postprocess", "hmvit_pillar_scatter",
)


class RowGemmArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("N", C.c_int32), ("n_out", C.c_int32),
                ("mode", C.c_void_p), ("record_len", C.c_void_p), ("ego_only", C.c_int32),
                ("a", C.c_void_p), ("w", C.c_void_p * 2), ("bias", C.c_void_p),
                ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_eps", C.c_float),
                ("resid", C.c_void_p), ("out", C.c_void_p), ("ln_stats", C.c_void_p)]


class ChainArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("N", C.c_int32), ("mode", C.c_void_p), ("record_len", C.c_void_p),
                ("ego_only", C.c_int32), ("o", C.c_void_p), ("resid", C.c_void_p), ("out", C.c_void_p),
                ("wa", C.c_void_p * 2), ("ba", C.c_void_p), ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p),
                ("ln_eps", C.c_float), ("w1", C.c_void_p * 2), ("b1", C.c_void_p), ("w2", C.c_void_p * 2), ("b2", C.c_void_p),
                ("stats_out", C.c_void_p)]


class HeadArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("N", C.c_int32), ("mode", C.c_void_p), ("record_len", C.c_void_p),
                ("x", C.c_void_p), ("w1", C.c_void_p * 2), ("b1", C.c_void_p), ("w2", C.c_void_p * 2), ("b2", C.c_void_p),
                ("out", C.c_void_p)]


class AttnArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("kind", C.c_int32), ("ego_only", C.c_int32), ("impl", C.c_int32), ("records_valid", C.c_int32),
                ("mode", C.c_void_p), ("record_len", C.c_void_p), ("cav_mask", C.c_void_p), ("T", C.c_void_p),
                ("cell", C.c_double), ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
                ("bk", C.c_void_p), ("bv", C.c_void_p), ("bias_table", C.c_void_p), ("key_mask", C.c_void_p),
                ("out", C.c_void_p), ("lse", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class WgradArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("N", C.c_int32), ("mode", C.c_void_p), ("record_len", C.c_void_p),
                ("ego_only", C.c_int32), ("a", C.c_void_p), ("a_rows_bf16", C.c_int32), ("b", C.c_void_p),
                ("b_rows_bf16", C.c_int32), ("b_stats", C.c_void_p), ("dw", C.c_void_p),
                ("dw_rows", C.c_int32), ("dw_row0", C.c_int32)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("kind", C.c_int32), ("ego_only", C.c_int32),
                ("mode", C.c_void_p), ("record_len", C.c_void_p), ("cav_mask", C.c_void_p), ("T", C.c_void_p),
                ("cell", C.c_double), ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
                ("bk", C.c_void_p), ("bv", C.c_void_p), ("bias_table", C.c_void_p),
                ("o", C.c_void_p), ("d_o", C.c_void_p), ("lse", C.c_void_p),
                ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
                ("dbk", C.c_void_p), ("dbv", C.c_void_p), ("dbias_table", C.c_void_p)]


class DecoderArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("num_convs", C.c_int32), ("anchor_number", C.c_int32),
                ("ego_mode", C.c_void_p), ("x", C.c_void_p), ("conv_w", C.c_void_p), ("conv_b", C.c_void_p),
                ("head_w", C.c_void_p), ("head_b", C.c_void_p), ("psm", C.c_void_p), ("rm", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class PostArgs(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("A", C.c_int32), ("psm", C.c_void_p), ("rm", C.c_void_p),
                ("anchor_box", C.c_void_p), ("transformation_matrix", C.c_void_p), ("order_hwl", C.c_int32),
                ("score_threshold", C.c_float), ("nms_thresh", C.c_float), ("range", C.c_float * 4),
                ("out_boxes", C.c_void_p), ("out_scores", C.c_void_p), ("out_count", C.c_void_p), ("status", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class PillarArgs(C.Structure):
    _fields_ = [("M", C.c_int32), ("P", C.c_int32), ("voxel_features", C.c_void_p), ("voxel_coords", C.c_void_p),
                ("voxel_num_points", C.c_void_p), ("w", C.c_void_p), ("b", C.c_void_p),
                ("voxel_size", C.c_float * 3), ("offset", C.c_float * 3),
                ("nx", C.c_int32), ("ny", C.c_int32), ("n_agents", C.c_int32), ("canvas", C.c_void_p), ("channels_last", C.c_int32)]


class StageWeights(C.Structure):
    _fields_ = [("wqkv", C.c_void_p * 2), ("bqkv", C.c_void_p), ("bk", C.c_void_p), ("bv", C.c_void_p),
                ("wa", C.c_void_p * 2), ("ba", C.c_void_p),
                ("ln1_g", C.c_void_p), ("ln1_b", C.c_void_p), ("ln2_g", C.c_void_p), ("ln2_b", C.c_void_p),
                ("w1", C.c_void_p * 2), ("b1", C.c_void_p), ("w2", C.c_void_p * 2), ("b2", C.c_void_p),
                ("bias_table", C.c_void_p), ("w1h", C.c_void_p * 2), ("w2h", C.c_void_p * 2)]


class FusionArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("L", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("num_iters", C.c_int32), ("head", C.c_int32), ("skip_dead", C.c_int32), ("unfused", C.c_int32),
                ("x", C.c_void_p), ("T", C.c_void_p), ("mode", C.c_void_p), ("record_len", C.c_void_p),
                ("cav_mask", C.c_void_p), ("cell", C.c_double), ("ln_eps", C.c_float),
                ("stage", StageWeights * 2),
                ("head_w1", C.c_void_p * 2), ("head_b1", C.c_void_p), ("head_w2", C.c_void_p * 2), ("head_b2", C.c_void_p),
                ("xres", C.c_void_p), ("workspace", C.c_void_p), ("out", C.c_void_p),
                ("head_w1h", C.c_void_p * 2), ("head_w2h", C.c_void_p * 2), ("attn_impl", C.c_int32)]


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run __graft_entry__.build() "
            "(nvcc -gencode arch=compute_100a,code=sm_100a). There is no CPU / eager fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.hmvit_abi_version.restype = C.c_int
    lib.hmvit_last_error.restype = C.c_char_p
    lib.hmvit_rowgemm.argtypes = [C.c_int, C.POINTER(RowGemmArgs), C.c_void_p]
    lib.hmvit_rowgemm.restype = C.c_int
    lib.hmvit_out_ffn_chain.argtypes = [C.POINTER(ChainArgs), C.c_void_p]
    lib.hmvit_out_ffn_chain.restype = C.c_int
    lib.hmvit_ffn_head.argtypes = [C.POINTER(HeadArgs), C.c_void_p]
    lib.hmvit_ffn_head.restype = C.c_int
    lib.hmvit_group_attn.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    lib.hmvit_group_attn.restype = C.c_int
    lib.hmvit_attn_records.argtypes = [C.POINTER(AttnArgs), C.c_void_p]
    lib.hmvit_attn_records.restype = C.c_int
    lib.hmvit_group_attn_workspace_bytes.argtypes = [C.c_int32] * 5
    lib.hmvit_group_attn_workspace_bytes.restype = C.c_size_t
    lib.hmvit_warp_bilinear.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_double, C.c_void_p]
    lib.hmvit_warp_bilinear.restype = C.c_int
    lib.hmvit_roi_cav_mask.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_double, C.c_void_p]
    lib.hmvit_roi_cav_mask.restype = C.c_int
    lib.hmvit_fusion_workspace_bytes.argtypes = [C.c_int32] * 6
    lib.hmvit_fusion_workspace_bytes.restype = C.c_size_t
    lib.hmvit_fusion_forward.argtypes = [C.POINTER(FusionArgs), C.c_void_p]
    lib.hmvit_fusion_forward.restype = C.c_int
    lib.hmvit_fusion_launch_count.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.hmvit_fusion_launch_count.restype = C.c_int
    i32, vp = C.c_int32, C.c_void_p
    lib.hmvit_bwd_row_stats.argtypes = [vp, vp, i32, i32, i32, vp, i32, C.c_float, vp]
    lib.hmvit_bwd_layernorm.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, vp, i32, vp]
    lib.hmvit_bwd_gelu.argtypes = [vp, vp, C.c_size_t, vp]
    lib.hmvit_bwd_cast_bf16.argtypes = [vp, vp, C.c_size_t, vp]
    lib.hmvit_bwd_colsum.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, vp, i32, vp]
    lib.hmvit_bwd_wgrad.argtypes = [C.POINTER(WgradArgs), vp]
    lib.hmvit_bwd_dgrad_cat.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp, vp]
    lib.hmvit_bwd_dgrad_cat.restype = C.c_int
    lib.hmvit_bwd_cast_colsum.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp]
    lib.hmvit_bwd_cast_colsum.restype = C.c_int
    lib.hmvit_dropout.argtypes = [vp, vp, vp, i32, i32, i32, vp, i32, C.c_uint64, C.c_uint32, C.c_float, vp]
    lib.hmvit_dropout.restype = C.c_int
    lib.hmvit_group_attn_bwd.argtypes = [C.POINTER(AttnBwdArgs), vp]
    lib.hmvit_decoder_workspace_bytes.argtypes = [i32, i32, i32]
    lib.hmvit_decoder_workspace_bytes.restype = C.c_size_t
    lib.hmvit_decoder_forward.argtypes = [C.POINTER(DecoderArgs), vp]
    lib.hmvit_decoder_forward.restype = C.c_int
    lib.hmvit_postprocess_workspace_bytes.argtypes = [i32, i32, i32]
    lib.hmvit_postprocess_workspace_bytes.restype = C.c_size_t
    lib.hmvit_postprocess.argtypes = [C.POINTER(PostArgs), vp]
    lib.hmvit_postprocess.restype = C.c_int
    lib.hmvit_pillar_scatter.argtypes = [C.POINTER(PillarArgs), vp]
    lib.hmvit_pillar_scatter.restype = C.c_int
    for fn in ("hmvit_bwd_row_stats", "hmvit_bwd_layernorm", "hmvit_bwd_gelu", "hmvit_bwd_cast_bf16", "hmvit_bwd_colsum",
               "hmvit_bwd_wgrad", "hmvit_group_attn_bwd"):
        getattr(lib, fn).restype = C.c_int
    if lib.hmvit_abi_version() != ABI_VERSION:
        raise ImportError("libhmvit_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


class HmvitError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().hmvit_last_error()
        raise HmvitError(msg.decode() if msg else f"hmvit error {rc}")
