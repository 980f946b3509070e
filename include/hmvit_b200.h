/* hmvit_b200 -- C ABI of the B200 (sm_100a) HM-ViT fusion hot path.
 *
 * The reference (XHwind/HM-ViT, a fork of OpenCOOD) has no C/FFI plugin interface: the boundary of
 * this path is the Python nn.Module `HeteroFusion`
 *   (opencood/models/bevformer_point_pillar_hetero.py:22-49, called at :122).
 * This header is the C-ABI that a drop-in replacement of that module binds (ctypes stub shown in
 * INTEGRATION.md; hm-vit_b200/_lib.py is the binding this repo ships).  Conventions:
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer including the workspace (hmvit_fusion_workspace_bytes);
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t): no host synchronisation, no
 *     allocation, CUDA-graph capturable;
 *   - returns 0 on success, non-zero on error; hmvit_last_error() gives the message of the last
 *     failing call on the calling thread.  No C++ exception crosses the boundary.
 *
 * Layouts.  N = H*W tokens per agent, C = 256 channels, L agent slots per scene.
 *   "cm"   : fp32 [agents][256][N]      channel-major == the module's (B, L, C, H, W) layout
 *   "rows" : bf16 [agents*N][256]       token-major rows (512 B per token)
 *   mode       : int32 [B*L]   0 = camera, 1 = LiDAR (padded slots 0)
 *   record_len : int32 [B]     valid agents per scene (>= 1)
 *   cav_mask   : int32 [B*L]   regroup() mask (fuse_utils.py:8-61)
 *   T          : fp32  [B][L][L][4][4]  pairwise_t_matrix, [b][i][j] maps agent i -> agent j
 */
#ifndef HMVIT_B200_H_
#define HMVIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMVIT_ABI_VERSION 9

/* error codes */
#define HMVIT_OK 0
#define HMVIT_ERR_ARG 1      /* bad argument (shape not supported, null pointer, ...) */
#define HMVIT_ERR_CUDA 2     /* a CUDA runtime / driver call failed */

int hmvit_abi_version(void);
const char* hmvit_last_error(void);

/* ---- typed row-GEMM (grouped by agent modality) -------------------------------------------------
 * Replaces the per-(b, agent) nn.Linear launches of
 *   HeteroAttention.to_qkv / to_out        opencood/models/sub_modules/hetero_fusion.py:111-152
 *   HeteroLayerNorm / HeteroFeedForward    opencood/models/base_transformer.py:138-192
 *   HeteroFusion.mlp_head                  opencood/models/bevformer_point_pillar_hetero.py:47-48 */
enum {
  HMVIT_GEMM_QKV = 0,   /* A = LN_type(x cm) bf16; out rows bf16 [n_chunks/2][B*L*N][256] = A W^T + bias     */
  HMVIT_GEMM_OUT = 1,   /* A = attention output rows bf16; out cm = A W^T + bias + resid cm                  */
  HMVIT_GEMM_FFN1 = 2,  /* A = LN_type(x cm) tf32; out cm = tf32(gelu(A W^T + bias))                         */
  HMVIT_GEMM_FFN2 = 3,  /* A = hidden cm tf32; out cm = A W^T + bias + resid cm                              */
  HMVIT_GEMM_HEAD1 = 4, /* A = x cm tf32 (slot 0 only, no norm); out cm = tf32(gelu(A W^T + bias))           */
  HMVIT_GEMM_HEAD2 = 5, /* A = hidden cm tf32 (slot 0 only); out [B][256][N] = A W^T + bias                  */
  HMVIT_GEMM_QKV_NOLN = 6, /* like QKV but A = x cm cast to bf16 without LayerNorm (unit-level attention API) */
  /* generic typed linears used by the backward pass (dgrad = a linear with the transposed weight) and by the
     recomputation of forward intermediates; all honour ego_only and write every slot's own rows */
  HMVIT_GEMM_LN_LIN_CM = 7,   /* A = LN_type(x cm) tf32; out cm = A W^T + bias                                */
  HMVIT_GEMM_LIN_CM = 8,      /* A = x cm tf32;          out cm = A W^T + bias                                */
  HMVIT_GEMM_LIN_ROWS = 9,    /* A = x cm tf32;          out rows bf16 [B*L*N][256] = A W^T + bias            */
  HMVIT_GEMM_ROWS_LIN_CM = 10 /* A = rows bf16;          out cm = A W^T + bias (+ resid cm when resid != NULL) */
};

typedef struct {
  int32_t B, L, N;            /* scenes, slots per scene, tokens per agent */
  int32_t n_out;              /* output channels of W (multiple of 128; 1280 for QKV, else 256) */
  const int32_t* mode;        /* [B*L] */
  const int32_t* record_len;  /* [B] */
  int32_t ego_only;           /* QKV: Q only for slot 0 and K/V only for the ego type of slot 0;
                                 other variants: process slot 0 tiles only */
  const void* a;              /* A operand: fp32 cm, or bf16 rows for HMVIT_GEMM_OUT */
  const void* w[2];           /* weights per modality type, row-major [n_out][256]: bf16 for QKV/OUT,
                                 fp32 holding tf32-rounded values for the FFN and HEAD variants */
  const float* bias;          /* [2][n_out] */
  const float* ln_gamma;      /* [2][256] (QKV, FFN1), or NULL when the affine is folded into w / bias */
  const float* ln_beta;       /* [2][256] or NULL */
  float ln_eps;
  const float* resid;         /* fp32 cm (OUT, FFN2) */
  void* out;                  /* see variant */
  const float* ln_stats;      /* QKV and LN_LIN_CM only, optional: [B*L][N][2] (mean, rstd) of every row of `a`, as written by
                                 hmvit_out_ffn_chain(stats_out); NULL = compute the statistics in the kernel */
} HmvitRowGemmArgs;

int hmvit_rowgemm(int variant, const HmvitRowGemmArgs* args, void* stream);

/* ---- fused output projection + residual + pre-norm FFN + residual --------------------------------------
 * One kernel for  x' = x + O W_a^T + b_a ;  x'' = x' + W_2 gelu(W_1 LN_type(x') + b_1) + b_2 , i.e.
 *   HeteroAttention.to_out + residual   opencood/models/sub_modules/hetero_fusion.py:142-152, 399, 442
 *   HeteroPreNormResidual(HeteroFeedForward)   opencood/models/base_transformer.py:129-136, 180-192
 * (same arithmetic as HMVIT_GEMM_OUT -> FFN1 -> FFN2, intermediates kept in TMEM / shared memory). */
typedef struct {
  int32_t B, L, N;
  const int32_t* mode;
  const int32_t* record_len;
  int32_t ego_only;           /* process slot 0 tiles only */
  const void* o;              /* attention output, bf16 rows [B*L*N][256] */
  const float* resid;         /* x, fp32 cm [B*L][256][N] */
  float* out;                 /* x'', fp32 cm [B*L][256][N]; may be the same buffer as resid */
  const void* wa[2];          /* bf16 [256][256] per type */
  const float* ba;            /* [2][256] */
  const float* ln_gamma;      /* [2][256], or NULL when the affine is folded into w1 / b1 */
  const float* ln_beta;       /* [2][256] or NULL */
  float ln_eps;
  const void* w1[2];          /* fp16 [256][256] (same 11-bit significand as tf32, twice the MMA rate) */
  const float* b1;            /* [2][256] */
  const void* w2[2];          /* fp16 [256][256] */
  const float* b2;            /* [2][256] */
  float* stats_out;           /* optional [B*L][N][2]: (mean, rstd) of every output row, the LayerNorm statistics
                                 the next stage's QKV projection needs; NULL = not written */
} HmvitChainArgs;

int hmvit_out_ffn_chain(const HmvitChainArgs* args, void* stream);

/* ---- typed feed-forward head on the ego rows ------------------------------------------------------
 * Replaces HeteroFusion.mlp_head applied to the ego slice
 *   opencood/models/bevformer_point_pillar_hetero.py:36, 46-48
 *   (HeteroFeedForward, opencood/models/base_transformer.py:180-192: Linear - GELU(erf) - Linear per agent type,
 *    no LayerNorm, no residual).
 * out[b] = W_2 gelu(W_1 x[b, slot 0] + b_1) + b_2, one launch of the chain kernel's head instance (fp16 operands, fp32 accumulate). */
typedef struct {
  int32_t B, L, N;
  const int32_t* mode;
  const int32_t* record_len;
  const float* x;             /* fp32 cm [B*L][256][N]; only slot 0 of every scene is read */
  const void* w1[2];          /* fp16 [256][256] per type */
  const float* b1;            /* [2][256] */
  const void* w2[2];          /* fp16 [256][256] per type */
  const float* b2;            /* [2][256] */
  float* out;                 /* fp32 cm [B][256][N] */
} HmvitHeadArgs;

int hmvit_ffn_head(const HmvitHeadArgs* args, void* stream);

/* ---- fused warp + mask + multi-agent window / grid attention ---------------------------------------
 * Replaces HeteroFusionBlock.warp_features + the ego loop around HeteroAttention.forward
 *   opencood/models/sub_modules/hetero_fusion.py:338-361, 373-397 (window) / 412-440 (grid), 187-277
 * and the warp / ROI-mask helpers it calls
 *   opencood/models/sub_modules/torch_transformation_utils.py:11-134, 254-355.
 * Three implementations of the same contract, selected explicitly (no environment switches):
 *   HMVIT_ATTN_FUSED   default: a key-record pass once per forward (fp64 source-pixel maps, bit-exact ROI visibility,
 *                      compaction of the visible keys; csrc/attn_fused.cuh tap_records_kernel) + ONE persistent
 *                      warp-specialised tcgen05 kernel per stage (csrc/attn_fa2.cuh fused_attn2_kernel): gather warps
 *                      blend the projected K' / V' taps straight into shared-memory operand tiles; S, P, D, Q and the
 *                      softmax denominator live in tensor memory; the relative position bias is part of the Q K^T
 *                      contraction.  Needs a workspace for the records; handles L <= 8 agents per scene and
 *                      B*L <= 1024 (larger shapes run HMVIT_ATTN_SINGLE).
 *   HMVIT_ATTN_SPLIT   round-1 form, kept as an independently written cross-check: warp + compaction pass writing dense
 *                      key / value tiles to the workspace, then an mma.sync dense attention pass (csrc/attn_split.cuh).
 *   HMVIT_ATTN_SINGLE  one mma.sync kernel per (ego, group, head group), gather inside (csrc/attn.cuh); no workspace,
 *                      any L. */
enum { HMVIT_ATTN_FUSED = 0, HMVIT_ATTN_SPLIT = 1, HMVIT_ATTN_SINGLE = 2 };

typedef struct {
  int32_t B, L, H, W;
  int32_t kind;               /* 0 = window partition, 1 = grid partition */
  int32_t ego_only;           /* only ego slot 0 of each scene */
  int32_t impl;               /* HMVIT_ATTN_* */
  int32_t records_valid;      /* HMVIT_ATTN_FUSED: 1 = `workspace` already holds this call's key records (same geometry
                                 arguments and kind; written by an earlier call) -- the record pass is skipped */
  const int32_t* mode;
  const int32_t* record_len;
  const int32_t* cav_mask;
  const float* T;             /* [B][L][L][16] */
  double cell;                /* voxel_size[0] * downsample_rate */
  const void* q;              /* bf16 rows [B*L*N][256]; W_q pre-multiplied by dim_head^-0.5 * log2(e) */
  const void* k;              /* bf16 rows [2 (ego type)][B*L*N][256], relation_att folded */
  const void* v;              /* bf16 rows [2 (ego type)][B*L*N][256], relation_msg folded */
  const float* bk;            /* [2 (ego type)][2 (source type)][256] folded key bias */
  const float* bv;            /* [2][2][256] folded value bias */
  const float* bias_table;    /* [225][8] relative_position_bias_table.weight */
  const uint8_t* key_mask;    /* optional [B*L][N] extra key mask indexed by (source slot, TARGET token); NULL = none.
                                 Used by the unit-level HeteroAttention.forward surface (explicit mask argument). */
  void* out;                  /* bf16 rows [B*L*N][256] */
  float* lse;                 /* optional [B*L*N][8]: log2-domain log-sum-exp per (query token, head), saved for
                                 hmvit_group_attn_bwd; NULL = not written (inference) */
  void* workspace;            /* scratch of >= hmvit_group_attn_workspace_bytes(impl, B, L, H, W) bytes, 256-byte aligned
                                 (FUSED: key records; SPLIT: compacted tiles; SINGLE: unused, may be NULL) */
  size_t workspace_bytes;
} HmvitAttnArgs;

int hmvit_group_attn(const HmvitAttnArgs* args, void* stream);
/* The key-record pass of HMVIT_ATTN_FUSED on its own (geometry fields, kind and workspace of `args`; q / k / v / out are not
 * read): the poses do not change between the block iterations, so a caller that runs several attention stages of one kind
 * computes the records once and passes records_valid = 1 afterwards (hmvit_fusion_forward does this for both kinds). */
int hmvit_attn_records(const HmvitAttnArgs* args, void* stream);
size_t hmvit_group_attn_workspace_bytes(int32_t impl, int32_t B, int32_t L, int32_t H, int32_t W);

/* ---- stand-alone spatial warp and ROI mask (unit-parity surface) ------------------------------------
 * SpatialTransformation.forward  opencood/models/sub_modules/spatial_transformation.py:16-44
 *   x, out: fp32 [n][C][H][W]; T: fp32 [n][16] (source -> target); bilinear, zero padding.
 * get_roi_and_cav_mask           opencood/models/sub_modules/torch_transformation_utils.py:11-49
 *   out: fp32 [B][H][W][1][L] = nearest-warp visibility * cav_mask. */
int hmvit_warp_bilinear(const float* x, const float* T, float* out, int32_t n, int32_t C, int32_t H, int32_t W,
                        double cell, void* stream);
int hmvit_roi_cav_mask(const float* T, const int32_t* cav_mask, float* out, int32_t B, int32_t L, int32_t H,
                       int32_t W, double cell, void* stream);

/* ---- whole fusion forward ----------------------------------------------------------------------------
 * HeteroFusion.forward            opencood/models/bevformer_point_pillar_hetero.py:39-49
 * HeteroFusionBlock.forward       opencood/models/sub_modules/hetero_fusion.py:446-458 (head == 0) */
typedef struct {
  const void* wqkv[2];        /* bf16 [1280][256] per source type: {Wq*scale*log2e, A(te=0)Wk, A(te=1)Wk, M(te=0)^T Wv, M(te=1)^T Wv} */
  const float* bqkv;          /* [2][1280] (K / V columns: zero, or W beta when the LayerNorm affine is folded) */
  const float* bk;            /* [2][2][256] */
  const float* bv;            /* [2][2][256] */
  const void* wa[2];          /* bf16 [256][256] */
  const float* ba;            /* [2][256] */
  const float* ln1_g; const float* ln1_b;   /* attention pre-norm  [2][256]; NULL when folded into wqkv / bqkv */
  const float* ln2_g; const float* ln2_b;   /* feed-forward pre-norm [2][256]; NULL when folded into w1 / b1 */
  const void* w1[2];          /* fp32 (tf32) [256][256] */
  const float* b1;            /* [2][256] */
  const void* w2[2];          /* fp32 (tf32) [256][256] */
  const float* b2;            /* [2][256] */
  const float* bias_table;    /* [225][8] */
  const void* w1h[2];         /* fp16 [256][256] copies of w1 / w2 for the fused chain kernel (w1 / w2 serve the unfused row-GEMMs) */
  const void* w2h[2];
} HmvitStageWeights;

typedef struct {
  int32_t B, L, H, W;
  int32_t num_iters;          /* block applications (weights shared) */
  int32_t head;               /* 1: ego slice + mlp_head -> out [B][256][N]; 0: stop after the blocks */
  int32_t skip_dead;          /* 1 (with head): last grid stage computes ego-0 queries only (exact) */
  int32_t unfused;            /* 1: run OUT / FFN1 / FFN2 as three row-GEMMs instead of the fused chain kernel */
  const float* x;             /* fp32 cm [B*L][256][N], not modified */
  const float* T;
  const int32_t* mode;
  const int32_t* record_len;
  const int32_t* cav_mask;
  double cell;
  float ln_eps;
  HmvitStageWeights stage[2]; /* [0] window, [1] grid */
  const void* head_w1[2];     /* fp32 (tf32) [256][256] */
  const float* head_b1;
  const void* head_w2[2];
  const float* head_b2;
  float* xres;                /* fp32 cm [B*L][256][N]: residual stream / block output (valid slots) */
  void* workspace;            /* hmvit_fusion_workspace_bytes(B, L, H, W, unfused, attn_impl) bytes, 256-byte aligned */
  float* out;                 /* fp32 [B][256][N] (head == 1) */
  const void* head_w1h[2];    /* fp16 [256][256] copies of head_w1 / head_w2 for the fused head kernel */
  const void* head_w2h[2];
  int32_t attn_impl;          /* HMVIT_ATTN_*: 0 = the fused persistent tcgen05 attention (default); the others are
                                 cross-check forms (shapes the fused kernel does not handle run HMVIT_ATTN_SINGLE) */
} HmvitFusionArgs;

size_t hmvit_fusion_workspace_bytes(int32_t B, int32_t L, int32_t H, int32_t W, int32_t unfused, int32_t attn_impl);
int hmvit_fusion_forward(const HmvitFusionArgs* args, void* stream);
/* number of kernel launches one hmvit_fusion_forward enqueues (for launch accounting): the key-record pass + per stage
 * {QKV, attention, chain}; head: 0 = no head, 1 = head as its own launch (skip_dead == 0), 2 = head with skip_dead (runs
 * inside the last stage's chain launch) */
int hmvit_fusion_launch_count(int32_t num_iters, int32_t head, int32_t attn_impl);

/* ---- backward pass (training configuration) ---------------------------------------------------------
 * The reference differentiates the fusion module with autograd (train_camera.py:172-193).  Here the
 * adjoint of the restructured forward is a set of kernels; the input-gradient GEMMs are hmvit_rowgemm
 * calls with transposed weights (variants 7-10), everything else is below.  Gradients are produced for
 * the FOLDED weights and mapped to the module parameters on the host through the folding function.
 * All calls are typed by mode[], skip padded slots and (ego_only != 0) every slot but 0. */

/* stats[a*N + tok] = (mean, rstd) over the 256 channels of x cm (biased variance, eps inside the sqrt) */
int hmvit_bwd_row_stats(const float* x, float* stats, int32_t B, int32_t L, int32_t N, const int32_t* record_len,
                        int32_t ego_only, float eps, void* stream);
/* LayerNorm (no affine) backward + residual: dx = dres + rstd (dz - mean(dz) - z mean(dz z)), z = (x - mean) rstd.
 * All cm fp32; dx may alias dres.  Adjoint of HeteroLayerNorm (base_transformer.py:171-177). */
int hmvit_bwd_layernorm(const float* dz, const float* x, const float* stats, const float* dres, float* dx, int32_t B,
                        int32_t L, int32_t N, const int32_t* record_len, int32_t ego_only, void* stream);
/* in place over n floats: hp <- gelu_erf(hp), dh <- dh * gelu_erf'(hp)   (nn.GELU, base_transformer.py:188) */
int hmvit_bwd_gelu(float* hp, float* dh, size_t n, void* stream);
/* dst[i] = bf16(src[i]), n floats (n % 4 == 0) */
int hmvit_bwd_cast_bf16(const float* src, void* dst, size_t n, void* stream);
/* dst = bf16(src) over the five gradient planes src [5][B*L*N][256] of the fused Q | K' | V' projection, and in the same pass
 * their typed column sums (the projection's bias gradient): db[type(agent of the row)*db_stride + p*256 + c] += src[p][row][c].
 * Rows of padded slots must be zero (hmvit_group_attn_bwd accumulates into a zero-filled buffer and never touches them). */
int hmvit_bwd_cast_colsum(const float* src, void* dst, float* db, int32_t db_stride, int32_t B, int32_t L, int32_t N,
                          const int32_t* mode, void* stream);
/* bias gradient: db[type*db_stride + c] += sum over tokens of y;  y is cm fp32 (rows_bf16 == 0) or bf16 rows */
int hmvit_bwd_colsum(const void* y, int32_t rows_bf16, float* db, int32_t db_stride, int32_t B, int32_t L, int32_t N,
                     const int32_t* mode, const int32_t* record_len, int32_t ego_only, void* stream);

/* Dropout of the training path (the shipped yaml trains with drop_out: 0.1):
 *   nn.Dropout after a_linears    opencood/models/sub_modules/hetero_fusion.py:66
 *   nn.Dropout x 2 in the FFN     opencood/models/base_transformer.py:186-190
 * out = resid + keep(seed, stream_id, i) * a / (1 - p) over cm fp32 tensors [B*L][256][N] (N % 4 == 0); resid may be NULL;
 * a == NULL exports the scaled mask itself (tests hand it to the CPU oracle).  keep() is Philox4x32-10 keyed by `seed`
 * with counter (element index / 4, stream_id): never stored, the backward regenerates it with the same arguments. */
int hmvit_dropout(const float* a, const float* resid, float* out, int32_t B, int32_t L, int32_t N, const int32_t* record_len,
                  int32_t ego_only, uint64_t seed, uint32_t stream_id, float p, void* stream);

/* typed weight gradient  dW[type][row0 + m][n] += sum_tok A(tok, m) B(tok, n)  (m, n < 256): operands rounded to bf16 while
 * staged, fp32 accumulate; tcgen05 (csrc/wgrad_tc.cuh) when at least one operand is cm, N % 64 == 0 and B*L <= 2048, else the
 * warp-level wmma kernel (csrc/bwd.cuh) */
typedef struct {
  int32_t B, L, N;
  const int32_t* mode;
  const int32_t* record_len;
  int32_t ego_only;
  const void* a;              /* output-gradient operand: cm fp32, or bf16 rows when a_rows_bf16 != 0 */
  int32_t a_rows_bf16;
  const void* b;              /* input-activation operand: cm fp32, or bf16 rows when b_rows_bf16 != 0 */
  int32_t b_rows_bf16;
  const float* b_stats;       /* optional (b cm only): [B*L*N][2] (mean, rstd) -- B is normalised on the fly */
  float* dw;                  /* [2][dw_rows][256] fp32, accumulated */
  int32_t dw_rows, dw_row0;
} HmvitWgradArgs;
int hmvit_bwd_wgrad(const HmvitWgradArgs* args, void* stream);

/* input gradient of the fused typed Q | K' | V' projection (adjoint of the GEMM of hmvit_rowgemm variant QKV;
 * autograd through HeteroAttention.to_qkv, opencood/models/sub_modules/hetero_fusion.py:111-132), ONE K = 1280 GEMM:
 *   out[a][c][tok] = sum_p sum_k dcat[p][a*N + tok][k] * w[type(a)][p*256 + c][k]
 * dcat: bf16 [5][B*L*N][256] (planes Q, K'|te=0, K'|te=1, V'|te=0, V'|te=1); w0 / w1: bf16 [1280][256], rows p*256.. hold
 * the TRANSPOSED plane p of the folded W_cat of type 0 / 1; out: cm fp32 [B*L][256][N], overwritten for every valid agent
 * (all valid agents: they are K/V sources even in the dead-query stage).  B*L <= 2048. */
int hmvit_bwd_dgrad_cat(const void* dcat, const void* w0, const void* w1, float* out, int32_t B, int32_t L, int32_t N,
                        const int32_t* mode, const int32_t* record_len, void* stream);

/* backward of hmvit_group_attn: same geometry arguments; q/k/v/o are the forward tensors, d_o the gradient of the
 * forward output, lse the statistics saved by the forward.  dq [R][256], dk / dv [2][R][256], dbk / dbv [2][2][256],
 * dbias_table [225][8]: fp32, ACCUMULATED (zero-fill before the first call).  R = B*L*N. */
typedef struct {
  int32_t B, L, H, W;
  int32_t kind, ego_only;
  const int32_t* mode;
  const int32_t* record_len;
  const int32_t* cav_mask;
  const float* T;
  double cell;
  const void* q; const void* k; const void* v;
  const float* bk; const float* bv; const float* bias_table;
  const void* o; const void* d_o;
  const float* lse;
  float* dq; float* dk; float* dv;
  float* dbk; float* dbv; float* dbias_table;
} HmvitAttnBwdArgs;
int hmvit_group_attn_bwd(const HmvitAttnBwdArgs* args, void* stream);

/* ---- detection decoder on the ego's fused feature --------------------------------
 * Replaces HeteroDecoder.forward(x, mode, use_upsample=False)
 *   opencood/models/sub_modules/hetero_decoder.py:42-74 (modality of the ego picks the camera / lidar weights),
 *   opencood/models/sub_modules/naive_decoder.py:63-92 (2 * num_layer x conv3x3 256->256 + BatchNorm + ReLU),
 *   the 1x1 cls / reg heads hetero_decoder.py:33-40, 62-70
 * as called from BevformerPointPillarHetero.forward (opencood/models/bevformer_point_pillar_hetero.py:125-126).
 * Eval-mode BatchNorm is folded into the convolution weights / bias by the caller (hm-vit_b200/decoder.py).
 * Kernels: csrc/decoder.cuh (TMA-shifted implicit-GEMM convolution on tcgen05, fp16 operands, fp32 accumulate). */
typedef struct {
  int32_t B, H, W;            /* scenes, BEV map; H, W divisible by 8 */
  int32_t num_convs;          /* 2 * num_layer convolutions */
  int32_t anchor_number;      /* A: psm has A channels, rm 7 A; 8 A <= 32 */
  const int32_t* ego_mode;    /* [B] 0 = camera, 1 = lidar (mode[:, 0]) */
  const float* x;             /* fp32 (B, 256, H, W): output of hmvit_fusion_forward */
  const void* conv_w;         /* fp16 [num_convs][2 (type)][9 (tap ky*3+kx)][256 out][256 in], BatchNorm folded */
  const float* conv_b;        /* fp32 [num_convs][2][256] folded bias */
  const float* head_w;        /* fp32 [2][8 A][256]: cls_head rows, then reg_head rows */
  const float* head_b;        /* fp32 [2][8 A] */
  float* psm;                 /* fp32 (B, A, H, W) */
  float* rm;                  /* fp32 (B, 7 A, H, W) */
  void* workspace;            /* hmvit_decoder_workspace_bytes(B, H, W), 1024-byte aligned */
  size_t workspace_bytes;
} HmvitDecoderArgs;
size_t hmvit_decoder_workspace_bytes(int32_t B, int32_t H, int32_t W);
int hmvit_decoder_forward(const HmvitDecoderArgs* args, void* stream);

/* ---- detection post-processing ----------------------------------------------------
 * Replaces VoxelPostprocessor.post_process for the intermediate-fusion case (one 'ego' output, batch size 1)
 *   opencood/data_utils/post_processor/voxel_postprocessor.py:232-343, 345-397 (delta_to_boxes3d)
 *   opencood/utils/box_utils.py:139-184, 258-296, 326-357, 575-620 (nms_rotated), 722-772
 * score threshold -> anchor decoding -> 8 corners -> projection -> sanity filters -> rotated NMS over the 1000 best-scored
 * candidates (IoU of the quadrilaterals spanned by corners 0..3, fp64) -> range mask.  Kernels: csrc/postproc.cuh.
 * Outputs stay on the device: out_boxes [1000][8][3], out_scores [1000], out_count [1] (boxes in the order the greedy NMS
 * picked them, i.e. by descending score), status [1] = 1 if more than 16384 anchors passed the threshold and filters. */
typedef struct {
  int32_t H, W, A;            /* feature map of psm / rm, anchors per cell */
  const float* psm;           /* fp32 (1, A, H, W) */
  const float* rm;            /* fp32 (1, 7 A, H, W) */
  const float* anchor_box;    /* fp32 (H, W, A, 7): x, y, z, h, w, l, r ('hwl') */
  const float* transformation_matrix;   /* fp32 4x4 row-major, or NULL ('no_post_projection') */
  int32_t order_hwl;          /* 1: params['order'] == 'hwl', 0: 'lwh' */
  float score_threshold, nms_thresh;
  float range[4];             /* GT_RANGE x_min, y_min, x_max, y_max */
  float* out_boxes; float* out_scores; int32_t* out_count; int32_t* status;
  void* workspace;            /* hmvit_postprocess_workspace_bytes(H, W, A), 256-byte aligned */
  size_t workspace_bytes;
} HmvitPostArgs;
size_t hmvit_postprocess_workspace_bytes(int32_t H, int32_t W, int32_t A);
int hmvit_postprocess(const HmvitPostArgs* args, void* stream);

/* ---- PointPillar front end (LiDAR branch in front of the fusion) ------------------------------
 * Replaces PillarVFE.forward (opencood/models/sub_modules/pillar_vfe.py:100-146) with ONE PFN layer
 * (PFNLayer.forward, :32-54: Linear(10 -> 64, no bias) + BatchNorm1d in eval mode + ReLU + max over the point slots)
 * for use_absolute_xyz = true, with_distance = false (the shipped yaml), and PointPillarScatter.forward
 * (opencood/models/sub_modules/point_pillar_scatter.py:15-48) in one kernel: voxels -> dense canvas.
 *   voxel_features  fp32 [M][P][4] (x, y, z, intensity; P <= 32 point slots, padded slots zero like the reference's input)
 *   voxel_coords    int32 [M][4]   (agent, z, y, x)         voxel_num_points  int32 [M] (>= 1)
 *   w               fp32 [64][10]  BatchNorm-folded weight  W * gamma / sqrt(var + eps)
 *   b               fp32 [64]      BatchNorm-folded shift   beta - mean * gamma / sqrt(var + eps)
 *   voxel_size / offset  fp32 [3]  (x, y, z): cell size and centre offset = size / 2 + range_min (pillar_vfe.py:85-90)
 *   canvas          fp32 (n_agents, 64, ny, nx) -- or [n_agents][ny][nx][64] with channels_last != 0 (the memory format cuDNN's
 *                   tensor-core convolutions use) --, overwritten (zero-filled, then the pillars' maxima)
 * All pointers are device pointers.  Padded point slots contribute relu(b) to the maximum, as in the reference (it masks
 * the features, not the outputs). */
typedef struct {
  int32_t M, P;
  const float* voxel_features; const int32_t* voxel_coords; const int32_t* voxel_num_points;
  const float* w; const float* b;
  float voxel_size[3]; float offset[3];
  int32_t nx, ny, n_agents;
  float* canvas;
  int32_t channels_last;
} HmvitPillarArgs;
int hmvit_pillar_scatter(const HmvitPillarArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HMVIT_B200_H_ */
