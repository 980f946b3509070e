// C-ABI of libhmvit_b200.so (declared in include/hmvit_b200.h): argument checking, TMA tensor-map
// encoding, kernel launches.  No torch types; everything is enqueued on the caller's stream.
#include "../../include/hmvit_b200.h"
#include "rowgemm.cuh"
#include "attn.cuh"
#include "attn_split.cuh"
#include "attn_fused.cuh"
#include "attn_fa2.cuh"
#include "attn_bwd.cuh"
#include "bwd.cuh"
#include "wgrad_tc.cuh"
#include "dgrad_cat.cuh"
#include "dropout.cuh"
#include "chain.cuh"
#include "qkv.cuh"
#include "decoder.cuh"
#include "postproc.cuh"
#include "pillar.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>

using namespace hmvit;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define HMVIT_CHECK_ARG(cond, msg) do { if (!(cond)) return fail(HMVIT_ERR_ARG, std::string("hmvit: ") + msg); } while (0)
#define HMVIT_CHECK_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
  return fail(HMVIT_ERR_CUDA, std::string("hmvit: " #expr ": ") + cudaGetErrorString(e_)); } while (0)

extern "C" int hmvit_abi_version(void) { return HMVIT_ABI_VERSION; }
extern "C" const char* hmvit_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// TMA tensor maps for the weight matrices ([n_out rows][256] row-major, box = 128 rows x 128 bytes)
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;
static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
}
static int make_weight_tmap(CUtensorMap* map, const void* w, long long n_out, int es, int box_rows = 128, bool f16 = false) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return fail(HMVIT_ERR_CUDA, "hmvit: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {256, static_cast<cuuint64_t>(n_out)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(256) * es};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, es == 2 ? (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                        const_cast<void*>(w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HMVIT_ERR_CUDA, "hmvit: cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")");
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// row-GEMM
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}

template <bool kLN>
static int launch_qkv(const CUtensorMap& m0, const CUtensorMap& m1, const QkvParams& p, cudaStream_t st) {

  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(qkv_kernel<kLN>, cudaFuncAttributeMaxDynamicSharedMemorySize, QkvCfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  const long long tiles = static_cast<long long>(p.B) * p.L * ((p.N + QkvCfg::BM - 1) / QkvCfg::BM);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  qkv_kernel<kLN><<<grid, QkvCfg::THREADS, QkvCfg::SMEM_BYTES, st>>>(m0, m1, p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

template <int ES, int PRO, int EPI>
static int launch_rowgemm(const CUtensorMap& m0, const CUtensorMap& m1, const RowGemmParams& p, cudaStream_t st) {
  using Cfg = RowGemmCfg<ES>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(rowgemm_kernel<ES, PRO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  dim3 grid((p.N + Cfg::BM - 1) / Cfg::BM, p.B * p.L);
  rowgemm_kernel<ES, PRO, EPI><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(m0, m1, p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_rowgemm(int variant, const HmvitRowGemmArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "rowgemm: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->N > 0, "rowgemm: B, L, N must be positive");
  HMVIT_CHECK_ARG(a->B * a->L <= 65535, "rowgemm: B*L exceeds grid limit");
  HMVIT_CHECK_ARG(a->n_out > 0 && a->n_out % 128 == 0 && a->n_out <= 32 * 128, "rowgemm: n_out must be a multiple of 128");
  HMVIT_CHECK_ARG(a->mode && a->record_len && a->a && a->w[0] && a->w[1] && a->bias && a->out, "rowgemm: null pointer");
  const bool tf32 = (variant >= HMVIT_GEMM_FFN1 && variant <= HMVIT_GEMM_HEAD2) ||
                    (variant >= HMVIT_GEMM_LN_LIN_CM && variant <= HMVIT_GEMM_LIN_ROWS);
  HMVIT_CHECK_ARG(variant >= HMVIT_GEMM_QKV && variant <= HMVIT_GEMM_ROWS_LIN_CM, "rowgemm: unknown variant");
  if (variant == HMVIT_GEMM_QKV || variant == HMVIT_GEMM_QKV_NOLN) HMVIT_CHECK_ARG(a->n_out == 1280, "rowgemm: QKV expects n_out == 1280");
  else HMVIT_CHECK_ARG(a->n_out == 256, "rowgemm: n_out must be 256 for this variant");
  HMVIT_CHECK_ARG((a->ln_gamma == nullptr) == (a->ln_beta == nullptr), "rowgemm: ln_gamma and ln_beta must both be set or both be null");
  if (variant == HMVIT_GEMM_OUT || variant == HMVIT_GEMM_FFN2) HMVIT_CHECK_ARG(a->resid != nullptr, "rowgemm: residual missing");

  CUtensorMap m0, m1;
  int rc = make_weight_tmap(&m0, a->w[0], a->n_out, tf32 ? 4 : 2); if (rc) return rc;
  rc = make_weight_tmap(&m1, a->w[1], a->n_out, tf32 ? 4 : 2); if (rc) return rc;

  RowGemmParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.N = a->N; p.n_chunks = a->n_out / 128;
  p.mode = a->mode; p.record_len = a->record_len;
  p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_eps = a->ln_eps;
  p.bias = a->bias; p.resid_cm = a->resid; p.out_L = a->L;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (variant) {
    case HMVIT_GEMM_QKV:
    case HMVIT_GEMM_QKV_NOLN: {
      QkvParams q;
      q.B = a->B; q.L = a->L; q.N = a->N; q.mode = a->mode; q.record_len = a->record_len; q.ego_only = a->ego_only ? 1 : 0;
      q.x_cm = static_cast<const float*>(a->a); q.ln_gamma = a->ln_gamma; q.ln_beta = a->ln_beta; q.ln_eps = a->ln_eps;
      q.stats_in = reinterpret_cast<const float2*>(a->ln_stats); q.bias = a->bias; q.out_rows = static_cast<__nv_bfloat16*>(a->out);
      return variant == HMVIT_GEMM_QKV ? launch_qkv<true>(m0, m1, q, st) : launch_qkv<false>(m0, m1, q, st);
    }
    case HMVIT_GEMM_OUT:
      p.a_rows = static_cast<const __nv_bfloat16*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      return launch_rowgemm<2, PRO_ROWS_BF16, EPI_CM_RESID>(m0, m1, p, st);
    case HMVIT_GEMM_FFN1:
      p.a_cm = static_cast<const float*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      return launch_rowgemm<4, PRO_CM_LN, EPI_CM_GELU>(m0, m1, p, st);
    case HMVIT_GEMM_FFN2:
      p.a_cm = static_cast<const float*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      return launch_rowgemm<4, PRO_CM_CAST, EPI_CM_RESID>(m0, m1, p, st);
    case HMVIT_GEMM_HEAD1:
      p.a_cm = static_cast<const float*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = 1;
      return launch_rowgemm<4, PRO_CM_CAST, EPI_CM_GELU>(m0, m1, p, st);
    case HMVIT_GEMM_HEAD2:
      p.a_cm = static_cast<const float*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = 1; p.out_L = 1;
      return launch_rowgemm<4, PRO_CM_CAST, EPI_CM_STORE>(m0, m1, p, st);
    case HMVIT_GEMM_LN_LIN_CM:
      p.a_cm = static_cast<const float*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.ln_stats = reinterpret_cast<const float2*>(a->ln_stats);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      return launch_rowgemm<4, PRO_CM_LN, EPI_CM_STORE>(m0, m1, p, st);
    case HMVIT_GEMM_LIN_CM:
      p.a_cm = static_cast<const float*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      return launch_rowgemm<4, PRO_CM_CAST, EPI_CM_STORE>(m0, m1, p, st);
    case HMVIT_GEMM_LIN_ROWS:
      p.a_cm = static_cast<const float*>(a->a); p.out_rows = static_cast<__nv_bfloat16*>(a->out);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      return launch_rowgemm<4, PRO_CM_CAST, EPI_ROWS_BF16>(m0, m1, p, st);
    case HMVIT_GEMM_ROWS_LIN_CM:
      p.a_rows = static_cast<const __nv_bfloat16*>(a->a); p.out_cm = static_cast<float*>(a->out);
      p.tile_ego_only = a->ego_only ? 1 : 0;
      if (a->resid != nullptr) return launch_rowgemm<2, PRO_ROWS_BF16, EPI_CM_RESID>(m0, m1, p, st);
      return launch_rowgemm<2, PRO_ROWS_BF16, EPI_CM_STORE>(m0, m1, p, st);
    default:
      return fail(HMVIT_ERR_ARG, "hmvit: rowgemm: unknown variant");
  }
}

// ------------------------------------------------------------------------------------------------
// fused output projection + FFN chain
// ------------------------------------------------------------------------------------------------
// `h` != NULL (internal, hmvit_fusion_forward only): the ego tiles of the last stage also run the feed-forward head in
// the same launch (chain_kernel<2>); the block output `a->out` is then not written.
static int launch_chain(const HmvitChainArgs* a, const HmvitHeadArgs* h, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "chain: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->N > 0, "chain: B, L, N must be positive");
  HMVIT_CHECK_ARG(a->mode && a->record_len && a->o && a->resid && a->out && a->wa[0] && a->wa[1] && a->ba && a->w1[0] && a->w1[1] && a->b1 && a->w2[0] && a->w2[1] && a->b2, "chain: null pointer");
  ChainMaps maps;
  int rc = make_weight_tmap(&maps.o, a->o, static_cast<long long>(a->B) * a->L * a->N, 2); if (rc) return rc;
  for (int t = 0; t < 2; ++t) {
    rc = make_weight_tmap(&maps.wa[t], a->wa[t], 256, 2, 256); if (rc) return rc;
    rc = make_weight_tmap(&maps.w1[t], a->w1[t], 256, 2, 256, true); if (rc) return rc;
    rc = make_weight_tmap(&maps.w2[t], a->w2[t], 256, 2, 256, true); if (rc) return rc;
    if (h != nullptr) {
      rc = make_weight_tmap(&maps.hw1[t], h->w1[t], 256, 2, 256, true); if (rc) return rc;
      rc = make_weight_tmap(&maps.hw2[t], h->w2[t], 256, 2, 256, true); if (rc) return rc;
    } else {
      maps.hw1[t] = maps.w1[t]; maps.hw2[t] = maps.w2[t];        // unused by the other instances
    }
  }
  ChainParams p;
  p.B = a->B; p.L = a->L; p.N = a->N; p.mode = a->mode; p.record_len = a->record_len; p.tile_ego_only = a->ego_only ? 1 : 0;
  p.resid_cm = a->resid; p.out_cm = a->out; p.ba = a->ba; p.ln_gamma = a->ln_gamma; p.ln_beta = a->ln_beta; p.ln_eps = a->ln_eps;
  p.b1 = a->b1; p.b2 = a->b2; p.stats_out = reinterpret_cast<float2*>(a->stats_out); p.out_L = a->L;
  p.hb1 = h ? h->b1 : nullptr; p.hb2 = h ? h->b2 : nullptr; p.head_out = h ? h->out : nullptr;
  if (h != nullptr) HMVIT_CHECK_ARG(a->ego_only && h->b1 && h->b2 && h->out, "chain: the fused head needs the ego-only stage and its biases / output");
  HMVIT_CHECK_ARG((a->ln_gamma == nullptr) == (a->ln_beta == nullptr), "chain: ln_gamma and ln_beta must both be set or both be null");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(chain_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  const long long tiles = static_cast<long long>(a->B) * a->L * ((a->N + ChainCfg::BM - 1) / ChainCfg::BM);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  if (h != nullptr) chain_kernel<2><<<grid, ChainCfg::THREADS, ChainCfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(maps, p);
  else chain_kernel<0><<<grid, ChainCfg::THREADS, ChainCfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(maps, p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_out_ffn_chain(const HmvitChainArgs* a, void* stream) { return launch_chain(a, nullptr, stream); }

// ------------------------------------------------------------------------------------------------
// typed feed-forward head on the ego rows (the chain kernel without projection, LayerNorm and residual)
// ------------------------------------------------------------------------------------------------
extern "C" int hmvit_ffn_head(const HmvitHeadArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "ffn_head: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->N > 0, "ffn_head: B, L, N must be positive");
  HMVIT_CHECK_ARG(a->mode && a->record_len && a->x && a->out && a->w1[0] && a->w1[1] && a->b1 && a->w2[0] && a->w2[1] && a->b2,
                  "ffn_head: null pointer");
  ChainMaps maps;
  memset(&maps, 0, sizeof(maps));
  for (int t = 0; t < 2; ++t) {
    int rc = make_weight_tmap(&maps.w1[t], a->w1[t], 256, 2, 256, true); if (rc) return rc;
    rc = make_weight_tmap(&maps.w2[t], a->w2[t], 256, 2, 256, true); if (rc) return rc;
  }
  ChainParams p;
  memset(&p, 0, sizeof(p));
  p.B = a->B; p.L = a->L; p.N = a->N; p.mode = a->mode; p.record_len = a->record_len; p.tile_ego_only = 1;
  p.resid_cm = a->x; p.out_cm = a->out; p.out_L = 1;
  p.ba = a->b1;                                  // unused by the head instance (staged with the other biases)
  p.b1 = a->b1; p.b2 = a->b2; p.ln_eps = 1e-5f;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ChainCfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  const long long tiles = static_cast<long long>(a->B) * a->L * ((a->N + ChainCfg::BM - 1) / ChainCfg::BM);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  chain_kernel<1><<<grid, ChainCfg::THREADS, ChainCfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(maps, p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// attention
// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// key records of ONE partition kind: [B*L][G][L*64] KeyRec + [B*L][G] visible-key counts
static size_t records_bytes(int B, int L, int H, int W) {
  const size_t G = static_cast<size_t>(H / 8) * (W / 8), BL = static_cast<size_t>(B) * L;
  return align_up(BL * G * L * kS * sizeof(KeyRec), 256) + align_up(BL * G * sizeof(int), 256);
}
// compacted key / value tiles of the split form (worst case: every key visible) + counts + key slots
static size_t split_bytes(int B, int L, int H, int W) {
  const size_t G = static_cast<size_t>(H / 8) * (W / 8), BL = static_cast<size_t>(B) * L;
  return 2 * BL * G * L * 2 * kBlobBytes + align_up(BL * G * 4, 256) + align_up(BL * G * L * 64, 256);
}
static bool fused_applicable(int B, int L) { return L <= kRecMaxL && B * L <= kFusedMaxAgents; }

extern "C" size_t hmvit_group_attn_workspace_bytes(int32_t impl, int32_t B, int32_t L, int32_t H, int32_t W) {
  if (B <= 0 || L <= 0 || H <= 0 || W <= 0) return 0;
  if (impl == HMVIT_ATTN_FUSED) return fused_applicable(B, L) ? records_bytes(B, L, H, W) : 0;
  if (impl == HMVIT_ATTN_SPLIT) return L <= kSplitMaxL ? split_bytes(B, L, H, W) : 0;
  return 0;
}

static int fill_attn_params(const HmvitAttnArgs* a, AttnParams& p) {
  HMVIT_CHECK_ARG(a != nullptr, "group_attn: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->H > 0 && a->W > 0, "group_attn: bad shape");
  HMVIT_CHECK_ARG(a->H % 8 == 0 && a->W % 8 == 0, "group_attn: H and W must be divisible by the window size 8");
  HMVIT_CHECK_ARG(a->B * a->L <= 65535, "group_attn: B*L exceeds grid limit");
  HMVIT_CHECK_ARG(a->kind == 0 || a->kind == 1, "group_attn: kind must be 0 (window) or 1 (grid)");
  HMVIT_CHECK_ARG(a->mode && a->record_len && a->cav_mask && a->T, "group_attn: null pointer");
  HMVIT_CHECK_ARG(a->cell > 0.0, "group_attn: cell size must be positive");
  p.B = a->B; p.L = a->L; p.H = a->H; p.W = a->W; p.kind = a->kind; p.ego_only = a->ego_only ? 1 : 0;
  p.mode = a->mode; p.record_len = a->record_len; p.cav_mask = a->cav_mask; p.T = a->T; p.cell = a->cell;
  p.q = static_cast<const __nv_bfloat16*>(a->q); p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v); p.bk = a->bk; p.bv = a->bv; p.bias_table = a->bias_table; p.key_mask = a->key_mask;
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.lse = a->lse;
  return HMVIT_OK;
}

// tap / visibility records of `nkinds` partition kinds starting at kind0, laid out kind after kind in `ws`
static int launch_records(const AttnParams& p, int kind0, int nkinds, void* ws, cudaStream_t st) {
  const size_t G = static_cast<size_t>(p.H / 8) * (p.W / 8), BL = static_cast<size_t>(p.B) * p.L;
  HMVIT_CHECK_ARG(p.L <= kRecMaxL, "attn_records: at most 8 agents per scene");
  RecParams rp;
  rp.a = p; rp.kind0 = kind0;
  // records of all kinds first, then the counts of all kinds (fused_ws_split() below mirrors this)
  rp.rec = static_cast<KeyRec*>(ws);
  rp.nvis = reinterpret_cast<int*>(static_cast<uint8_t*>(ws) + nkinds * align_up(BL * G * p.L * kS * sizeof(KeyRec), 256));
  dim3 grid(static_cast<unsigned>(G), static_cast<unsigned>(BL), nkinds);
  tap_records_kernel<<<grid, 256, 0, st>>>(rp);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}
static void fused_ws_split(const AttnParams& p, int nkinds, int which, void* ws, const KeyRec** rec, const int** nvis) {
  const size_t G = static_cast<size_t>(p.H / 8) * (p.W / 8), BL = static_cast<size_t>(p.B) * p.L;
  const size_t per_kind = BL * G * p.L * kS;
  *rec = static_cast<const KeyRec*>(ws) + which * per_kind;
  *nvis = reinterpret_cast<const int*>(static_cast<uint8_t*>(ws) + nkinds * align_up(per_kind * sizeof(KeyRec), 256)) + which * BL * G;
}

static int launch_fused_attn(const AttnParams& p, const KeyRec* rec, const int* nvis, cudaStream_t st) {
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(fused_attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa2Cfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  FusedAttnParams fp;
  fp.a = p; fp.rec = rec; fp.nvis = nvis;
  const long long items = static_cast<long long>(p.B) * (p.ego_only ? 1 : p.L) * (p.H / 8) * (p.W / 8) * 2;
  const long long cap = num_sms() & ~1;                        // one persistent CTA per SM, one head group each
  const int grid = static_cast<int>(items < cap ? items : cap);
  fused_attn2_kernel<<<grid, Fa2Cfg::THREADS, Fa2Cfg::SMEM_BYTES, st>>>(fp);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_attn_records(const HmvitAttnArgs* a, void* stream) {
  AttnParams p;
  int rc = fill_attn_params(a, p); if (rc) return rc;
  HMVIT_CHECK_ARG(fused_applicable(a->B, a->L), "attn_records: at most 8 agents per scene and B*L <= 1024");
  HMVIT_CHECK_ARG(a->workspace != nullptr && a->workspace_bytes >= records_bytes(a->B, a->L, a->H, a->W), "attn_records: workspace too small");
  HMVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "attn_records: workspace must be 256-byte aligned");
  return launch_records(p, a->kind, 1, a->workspace, static_cast<cudaStream_t>(stream));
}

extern "C" int hmvit_group_attn(const HmvitAttnArgs* a, void* stream) {
  AttnParams p;
  int rc = fill_attn_params(a, p); if (rc) return rc;
  HMVIT_CHECK_ARG(a->q && a->k && a->v && a->bk && a->bv && a->bias_table && a->out, "group_attn: null pointer");
  HMVIT_CHECK_ARG(a->impl >= HMVIT_ATTN_FUSED && a->impl <= HMVIT_ATTN_SINGLE, "group_attn: unknown impl");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(group_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(dense_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmem);
  });
  HMVIT_CHECK_CUDA(attr_err);
  int impl = a->impl;
  if (impl == HMVIT_ATTN_FUSED && !fused_applicable(a->B, a->L)) impl = HMVIT_ATTN_SINGLE;   // > 8 agents per scene
  if (impl != HMVIT_ATTN_SINGLE) {
    HMVIT_CHECK_ARG(a->workspace != nullptr, "group_attn: this implementation needs a workspace (hmvit_group_attn_workspace_bytes)");
    HMVIT_CHECK_ARG(a->workspace_bytes >= hmvit_group_attn_workspace_bytes(impl, a->B, a->L, a->H, a->W), "group_attn: workspace too small");
    HMVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "group_attn: workspace must be 256-byte aligned");
  }
  if (impl == HMVIT_ATTN_FUSED) {
    // persistent tcgen05 kernel fed by the key records of this partition kind (csrc/attn_fused.cuh)
    if (!a->records_valid) { rc = launch_records(p, a->kind, 1, a->workspace, st); if (rc) return rc; }
    const KeyRec* rec; const int* nvis;
    fused_ws_split(p, 1, 0, a->workspace, &rec, &nvis);
    return launch_fused_attn(p, rec, nvis, st);
  }
  dim3 grid((a->H / 8) * (a->W / 8) * 2, a->B * a->L);      // x: (token group, head group)
  if (impl == HMVIT_ATTN_SPLIT) {
    // warp + compaction pass, then mma.sync dense attention over the compacted key tiles (csrc/attn_split.cuh);
    // kept as an independently written cross-check of the fused kernel
    HMVIT_CHECK_ARG(a->L <= kSplitMaxL, "group_attn: the split form handles at most 8 agents per scene");
    const size_t G = static_cast<size_t>(a->H / 8) * (a->W / 8), BL = static_cast<size_t>(a->B) * a->L;
    SplitParams sp;
    sp.a = p;
    uint8_t* ws = static_cast<uint8_t*>(a->workspace);
    const size_t tile_bytes = BL * G * a->L * 2 * kBlobBytes;
    sp.kc = ws; ws += tile_bytes;
    sp.vc = ws; ws += tile_bytes;
    sp.nvis = reinterpret_cast<int*>(ws); ws += (BL * G * 4 + 255) / 256 * 256;
    sp.slots = ws;
    sp.tc_layout = 0;
    dim3 grid_c(static_cast<unsigned>(G), a->B * a->L);
    warp_compact_kernel<<<grid_c, kCompactThreads, 0, st>>>(sp);
    HMVIT_CHECK_CUDA(cudaGetLastError());
    dense_attn_kernel<<<grid, kDenseThreads, kDenseSmem, st>>>(sp);
    HMVIT_CHECK_CUDA(cudaGetLastError());
    return HMVIT_OK;
  }
  group_attn_kernel<<<grid, kAttnThreads, kAttnSmem, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// stand-alone warp / ROI mask (NCHW fp32, the reference's own layout)
// ------------------------------------------------------------------------------------------------
// One thread per output pixel (u fastest -> coalesced stores), channel loop inside; the four tap
// offsets / weights are computed once per pixel.
__global__ void __launch_bounds__(256) warp_bilinear_kernel(const float* __restrict__ x, const float* __restrict__ T,
                                                            float* __restrict__ out, int C, int H, int W, double cell,
                                                            int c_per_block) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (pix >= HW) return;
  const int v = pix / W, u = pix - v * W;
  const WarpMap wm = make_warp_map(T + static_cast<size_t>(n) * 16, H, W, cell);
  double sx, sy; warp_src(wm, u, v, sx, sy);
  const Taps tp = make_taps(sx, sy, H, W);
  const int i00 = tp.y0 * W + tp.x0;
  const int c0 = blockIdx.z * c_per_block, c1 = min(C, c0 + c_per_block);
  const float* src = x + (static_cast<size_t>(n) * C + c0) * HW;
  float* dst = out + (static_cast<size_t>(n) * C + c0) * HW + pix;
  for (int c = c0; c < c1; ++c, src += HW, dst += HW) {
    float acc = 0.f;
    if (tp.w00 != 0.f) acc += tp.w00 * __ldg(src + i00);
    if (tp.w01 != 0.f) acc += tp.w01 * __ldg(src + i00 + 1);
    if (tp.w10 != 0.f) acc += tp.w10 * __ldg(src + i00 + W);
    if (tp.w11 != 0.f) acc += tp.w11 * __ldg(src + i00 + W + 1);
    *dst = acc;
  }
}

extern "C" int hmvit_warp_bilinear(const float* x, const float* T, float* out, int32_t n, int32_t C, int32_t H, int32_t W,
                                   double cell, void* stream) {
  HMVIT_CHECK_ARG(x && T && out, "warp_bilinear: null pointer");
  HMVIT_CHECK_ARG(n > 0 && C > 0 && H > 0 && W > 0 && n <= 65535, "warp_bilinear: bad shape");
  HMVIT_CHECK_ARG(cell > 0.0, "warp_bilinear: cell size must be positive");
  const int cpb = 32;
  dim3 grid((H * W + 255) / 256, n, (C + cpb - 1) / cpb);
  warp_bilinear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, T, out, C, H, W, cell, cpb);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

__global__ void __launch_bounds__(256) roi_cav_mask_kernel(const float* __restrict__ T, const int* __restrict__ cav_mask,
                                                           float* __restrict__ out, int L, int H, int W, double cell) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= H * W) return;
  const int v = pix / W, u = pix - v * W;
  float* dst = out + (static_cast<size_t>(b) * H * W + pix) * L;
  for (int l = 0; l < L; ++l) {
    const WarpMap wm = make_warp_map(T + (static_cast<size_t>(b) * L + l) * 16, H, W, cell);
    double sx, sy; warp_src(wm, u, v, sx, sy);
    dst[l] = (warp_visible(sx, sy, H, W) && cav_mask[b * L + l] != 0) ? 1.0f : 0.0f;
  }
}

extern "C" int hmvit_roi_cav_mask(const float* T, const int32_t* cav_mask, float* out, int32_t B, int32_t L, int32_t H,
                                  int32_t W, double cell, void* stream) {
  HMVIT_CHECK_ARG(T && cav_mask && out, "roi_cav_mask: null pointer");
  HMVIT_CHECK_ARG(B > 0 && L > 0 && H > 0 && W > 0 && B <= 65535, "roi_cav_mask: bad shape");
  HMVIT_CHECK_ARG(cell > 0.0, "roi_cav_mask: cell size must be positive");
  dim3 grid((H * W + 255) / 256, B);
  roi_cav_mask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(T, cav_mask, out, L, H, W, cell);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// whole forward
// ------------------------------------------------------------------------------------------------
static int effective_attn_impl(int impl, int B, int L) {
  if (impl == HMVIT_ATTN_FUSED && !fused_applicable(B, L)) return HMVIT_ATTN_SINGLE;
  if (impl == HMVIT_ATTN_SPLIT && L > kSplitMaxL) return HMVIT_ATTN_SINGLE;
  return impl;
}

extern "C" size_t hmvit_fusion_workspace_bytes(int32_t B, int32_t L, int32_t H, int32_t W, int32_t unfused, int32_t attn_impl) {
  if (B <= 0 || L <= 0 || H <= 0 || W <= 0) return 0;
  const int impl = effective_attn_impl(attn_impl, B, L);
  const size_t rows = static_cast<size_t>(B) * L * H * W;
  size_t bytes = 0;
  bytes += align_up(rows * 256 * 2 * 5, 1024);   // q, k|te0, k|te1, v|te0, v|te1 (bf16 rows)
  bytes += align_up(rows * 256 * 2, 1024);       // attention output (bf16 rows)
  bytes += align_up(rows * 2 * 4, 1024);         // per-row LayerNorm statistics handed from one stage to the next
  if (unfused) bytes += align_up(rows * 256 * 4, 1024);   // FFN hidden (fp32 cm, tf32 values): unfused cross-check path only
  if (impl == HMVIT_ATTN_FUSED) bytes += align_up(2 * records_bytes(B, L, H, W), 1024);   // key records of both partition kinds
  if (impl == HMVIT_ATTN_SPLIT) bytes += align_up(split_bytes(B, L, H, W), 1024);         // compacted key / value tiles
  return bytes;
}

static bool fuse_head(const HmvitFusionArgs* a) {
  return a->head && !a->unfused && a->skip_dead && a->head_w1h[0] && a->head_w1h[1] && a->head_w2h[0] &&
         a->head_w2h[1] && a->head_b1 && a->head_b2 && a->out;
}

/* head: 0 = no head, 1 = head as its own launch, 2 = head with skip_dead (fused into the last stage's chain launch) */
extern "C" int hmvit_fusion_launch_count(int32_t num_iters, int32_t head, int32_t attn_impl) {
  const int head_launch = head == 1 ? 1 : 0;
  if (attn_impl == HMVIT_ATTN_FUSED) return 1 + num_iters * 2 * 3 + head_launch;   // key records (both kinds) + per stage {QKV, attention, chain}
  return num_iters * 2 * (attn_impl == HMVIT_ATTN_SPLIT ? 4 : 3) + head_launch;
}

extern "C" int hmvit_fusion_forward(const HmvitFusionArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "fusion_forward: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->H > 0 && a->W > 0, "fusion_forward: bad shape");
  HMVIT_CHECK_ARG(a->H % 8 == 0 && a->W % 8 == 0, "fusion_forward: H and W must be divisible by the window size 8");
  HMVIT_CHECK_ARG(a->num_iters >= 1, "fusion_forward: num_iters must be >= 1");
  HMVIT_CHECK_ARG(a->x && a->T && a->mode && a->record_len && a->cav_mask && a->xres && a->workspace, "fusion_forward: null pointer");
  HMVIT_CHECK_ARG(!a->head || a->out, "fusion_forward: out is null");
  HMVIT_CHECK_ARG(a->cell > 0.0, "fusion_forward: cell size must be positive");
  HMVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "fusion_forward: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int N = a->H * a->W;
  const size_t rows = static_cast<size_t>(a->B) * a->L * N;
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align_up(rows * 256 * 2 * 5, 1024);
  __nv_bfloat16* att = reinterpret_cast<__nv_bfloat16*>(ws);
  ws += align_up(rows * 256 * 2, 1024);
  float* stats = reinterpret_cast<float*>(ws);
  ws += align_up(rows * 2 * 4, 1024);
  float* hid = nullptr;
  if (a->unfused) { hid = reinterpret_cast<float*>(ws); ws += align_up(rows * 256 * 4, 1024); }
  HMVIT_CHECK_ARG(a->attn_impl >= HMVIT_ATTN_FUSED && a->attn_impl <= HMVIT_ATTN_SINGLE, "fusion_forward: unknown attn_impl");
  const int attn_impl = effective_attn_impl(a->attn_impl, a->B, a->L);
  const bool fused_attn = attn_impl == HMVIT_ATTN_FUSED;
  void* rec_ws = ws;                              // key records (fused) or compacted tiles (split)
  bool have_stats = false;                        // stats describe the rows currently in xres
  bool head_done = false;                         // the head ran inside the last stage's chain launch

  AttnParams geo;
  memset(&geo, 0, sizeof(geo));
  geo.B = a->B; geo.L = a->L; geo.H = a->H; geo.W = a->W; geo.mode = a->mode; geo.record_len = a->record_len;
  geo.cav_mask = a->cav_mask; geo.T = a->T; geo.cell = a->cell;
  if (fused_attn) {
    // the poses do not change between the block iterations: key records of both partition kinds, once per forward
    int rc = launch_records(geo, 0, 2, rec_ws, st); if (rc) return rc;
  }

  for (int it = 0; it < a->num_iters; ++it) {
    for (int kind = 0; kind < 2; ++kind) {
      const HmvitStageWeights& w = a->stage[kind];
      const bool first = (it == 0 && kind == 0);
      const int dead = (a->head && a->skip_dead && it == a->num_iters - 1 && kind == 1) ? 1 : 0;
      const float* xsrc = first ? a->x : a->xres;
      HmvitRowGemmArgs g;
      memset(&g, 0, sizeof(g));
      g.B = a->B; g.L = a->L; g.N = N; g.mode = a->mode; g.record_len = a->record_len; g.ego_only = dead; g.ln_eps = a->ln_eps;
      // typed LayerNorm + Q / K' / V' projections
      g.n_out = 1280; g.a = xsrc; g.w[0] = w.wqkv[0]; g.w[1] = w.wqkv[1]; g.bias = w.bqkv;
      g.ln_gamma = w.ln1_g; g.ln_beta = w.ln1_b; g.out = qkv; g.ln_stats = have_stats ? stats : nullptr;
      int rc = hmvit_rowgemm(HMVIT_GEMM_QKV, &g, stream); if (rc) return rc;
      g.ln_stats = nullptr;
      // warp + mask + attention
      HMVIT_CHECK_ARG(w.bk && w.bv && w.bias_table, "fusion_forward: attention biases missing");
      AttnParams p = geo;
      p.kind = kind; p.ego_only = dead;
      p.q = qkv; p.k = qkv + rows * 256; p.v = qkv + rows * 256 * 3; p.bk = w.bk; p.bv = w.bv; p.bias_table = w.bias_table;
      p.out = att;
      if (fused_attn) {
        const KeyRec* rec; const int* nvis;
        fused_ws_split(geo, 2, kind, rec_ws, &rec, &nvis);
        rc = launch_fused_attn(p, rec, nvis, st); if (rc) return rc;
      } else {
        // split / single cross-check forms, and shapes with more than 8 agents per scene (csrc/attn_split.cuh, attn.cuh)
        HmvitAttnArgs t;
        memset(&t, 0, sizeof(t));
        t.B = a->B; t.L = a->L; t.H = a->H; t.W = a->W; t.kind = kind; t.ego_only = dead; t.impl = attn_impl;
        if (attn_impl == HMVIT_ATTN_SPLIT) { t.workspace = rec_ws; t.workspace_bytes = split_bytes(a->B, a->L, a->H, a->W); }
        t.mode = a->mode; t.record_len = a->record_len; t.cav_mask = a->cav_mask; t.T = a->T; t.cell = a->cell;
        t.q = p.q; t.k = p.k; t.v = p.v; t.bk = w.bk; t.bv = w.bv; t.bias_table = w.bias_table; t.out = att;
        rc = hmvit_group_attn(&t, stream); if (rc) return rc;
      }
      if (a->unfused) {
        // output projection + residual
        g.n_out = 256; g.a = att; g.w[0] = w.wa[0]; g.w[1] = w.wa[1]; g.bias = w.ba; g.resid = xsrc; g.out = a->xres;
        rc = hmvit_rowgemm(HMVIT_GEMM_OUT, &g, stream); if (rc) return rc;
        // pre-norm feed-forward + residual
        g.a = a->xres; g.w[0] = w.w1[0]; g.w[1] = w.w1[1]; g.bias = w.b1; g.ln_gamma = w.ln2_g; g.ln_beta = w.ln2_b; g.out = hid;
        rc = hmvit_rowgemm(HMVIT_GEMM_FFN1, &g, stream); if (rc) return rc;
        g.a = hid; g.w[0] = w.w2[0]; g.w[1] = w.w2[1]; g.bias = w.b2; g.resid = a->xres; g.out = a->xres;
        rc = hmvit_rowgemm(HMVIT_GEMM_FFN2, &g, stream); if (rc) return rc;
        have_stats = false;
      } else {
        // output projection + residual + pre-norm feed-forward + residual, one kernel
        HmvitChainArgs c;
        memset(&c, 0, sizeof(c));
        c.B = a->B; c.L = a->L; c.N = N; c.mode = a->mode; c.record_len = a->record_len; c.ego_only = dead;
        c.o = att; c.resid = xsrc; c.out = a->xres;
        c.wa[0] = w.wa[0]; c.wa[1] = w.wa[1]; c.ba = w.ba; c.ln_gamma = w.ln2_g; c.ln_beta = w.ln2_b; c.ln_eps = a->ln_eps;
        HMVIT_CHECK_ARG(w.w1h[0] && w.w1h[1] && w.w2h[0] && w.w2h[1], "fusion_forward: fp16 feed-forward weights (w1h / w2h) missing");
        c.w1[0] = w.w1h[0]; c.w1[1] = w.w1h[1]; c.b1 = w.b1; c.w2[0] = w.w2h[0]; c.w2[1] = w.w2h[1]; c.b2 = w.b2;
        c.stats_out = stats;
        const bool last = (it == a->num_iters - 1 && kind == 1);
        if (last && dead && fuse_head(a)) {
          // last stage on the ego tiles only: the head runs inside the same launch
          HmvitHeadArgs h;
          memset(&h, 0, sizeof(h));
          h.B = a->B; h.L = a->L; h.N = N; h.mode = a->mode; h.record_len = a->record_len; h.x = a->xres;
          h.w1[0] = a->head_w1h[0]; h.w1[1] = a->head_w1h[1]; h.b1 = a->head_b1;
          h.w2[0] = a->head_w2h[0]; h.w2[1] = a->head_w2h[1]; h.b2 = a->head_b2; h.out = a->out;
          rc = launch_chain(&c, &h, stream); if (rc) return rc;
          head_done = true;
        } else {
          rc = launch_chain(&c, nullptr, stream); if (rc) return rc;
        }
        have_stats = true;
      }
    }
  }
  if (a->head && !head_done) {
    HMVIT_CHECK_ARG(a->head_b1 && a->head_b2, "fusion_forward: head biases missing");
    if (a->unfused) {
      HMVIT_CHECK_ARG(a->head_w1[0] && a->head_w1[1] && a->head_w2[0] && a->head_w2[1], "fusion_forward: head weights missing");
      HmvitRowGemmArgs g;
      memset(&g, 0, sizeof(g));
      g.B = a->B; g.L = a->L; g.N = N; g.mode = a->mode; g.record_len = a->record_len; g.ego_only = 1; g.ln_eps = a->ln_eps;
      g.n_out = 256; g.a = a->xres; g.w[0] = a->head_w1[0]; g.w[1] = a->head_w1[1]; g.bias = a->head_b1; g.out = hid;
      int rc = hmvit_rowgemm(HMVIT_GEMM_HEAD1, &g, stream); if (rc) return rc;
      g.a = hid; g.w[0] = a->head_w2[0]; g.w[1] = a->head_w2[1]; g.bias = a->head_b2; g.out = a->out;
      rc = hmvit_rowgemm(HMVIT_GEMM_HEAD2, &g, stream); if (rc) return rc;
    } else {
      HmvitHeadArgs h;
      memset(&h, 0, sizeof(h));
      h.B = a->B; h.L = a->L; h.N = N; h.mode = a->mode; h.record_len = a->record_len; h.x = a->xres;
      HMVIT_CHECK_ARG(a->head_w1h[0] && a->head_w1h[1] && a->head_w2h[0] && a->head_w2h[1], "fusion_forward: fp16 head weights missing");
      h.w1[0] = a->head_w1h[0]; h.w1[1] = a->head_w1h[1]; h.b1 = a->head_b1;
      h.w2[0] = a->head_w2h[0]; h.w2[1] = a->head_w2h[1]; h.b2 = a->head_b2; h.out = a->out;
      int rc = hmvit_ffn_head(&h, stream); if (rc) return rc;
    }
  }
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// backward pass
// ------------------------------------------------------------------------------------------------
extern "C" int hmvit_bwd_row_stats(const float* x, float* stats, int32_t B, int32_t L, int32_t N, const int32_t* record_len,
                                   int32_t ego_only, float eps, void* stream) {
  HMVIT_CHECK_ARG(x && stats && record_len, "bwd_row_stats: null pointer");
  HMVIT_CHECK_ARG(B > 0 && L > 0 && N > 0 && B * L <= 65535, "bwd_row_stats: bad shape");
  dim3 grid((N + 127) / 128, B * L);
  row_stats_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<float2*>(stats), L, N, record_len,
                                                                        ego_only ? 1 : 0, eps);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_bwd_layernorm(const float* dz, const float* x, const float* stats, const float* dres, float* dx, int32_t B,
                                   int32_t L, int32_t N, const int32_t* record_len, int32_t ego_only, void* stream) {
  HMVIT_CHECK_ARG(dz && x && stats && dres && dx && record_len, "bwd_layernorm: null pointer");
  HMVIT_CHECK_ARG(B > 0 && L > 0 && N > 0 && B * L <= 65535, "bwd_layernorm: bad shape");
  dim3 grid((N + 127) / 128, B * L);
  ln_bwd_cm_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(dz, x, reinterpret_cast<const float2*>(stats), dres, dx, L, N,
                                                                        record_len, ego_only ? 1 : 0);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

static int flat_grid(size_t n4) {
  const size_t blocks = (n4 + 255) / 256;
  const size_t cap = static_cast<size_t>(num_sms()) * 16;
  return static_cast<int>(blocks < cap ? (blocks ? blocks : 1) : cap);
}

extern "C" int hmvit_bwd_gelu(float* hp, float* dh, size_t n, void* stream) {
  HMVIT_CHECK_ARG(hp && dh, "bwd_gelu: null pointer");
  HMVIT_CHECK_ARG(n % 4 == 0, "bwd_gelu: n must be a multiple of 4");
  if (n == 0) return HMVIT_OK;
  gelu_bwd_kernel<<<flat_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4*>(hp), reinterpret_cast<float4*>(dh), n / 4);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_bwd_cast_bf16(const float* src, void* dst, size_t n, void* stream) {
  HMVIT_CHECK_ARG(src && dst, "bwd_cast_bf16: null pointer");
  HMVIT_CHECK_ARG(n % 4 == 0, "bwd_cast_bf16: n must be a multiple of 4");
  if (n == 0) return HMVIT_OK;
  cast_bf16_kernel<<<flat_grid(n / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n / 4);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_bwd_colsum(const void* y, int32_t rows_bf16, float* db, int32_t db_stride, int32_t B, int32_t L, int32_t N,
                                const int32_t* mode, const int32_t* record_len, int32_t ego_only, void* stream) {
  HMVIT_CHECK_ARG(y && db && mode && record_len, "bwd_colsum: null pointer");
  HMVIT_CHECK_ARG(B > 0 && L > 0 && N > 0 && B * L <= 65535 && db_stride >= 256, "bwd_colsum: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows_bf16) {
    dim3 grid((N + 255) / 256, B * L);
    colsum_rows_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(y), db, db_stride, L, N, mode, record_len, ego_only ? 1 : 0);
  } else {
    dim3 grid(256 / 8, B * L);
    colsum_cm_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(y), db, db_stride, L, N, mode, record_len, ego_only ? 1 : 0);
  }
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_dropout(const float* a, const float* resid, float* out, int32_t B, int32_t L, int32_t N,
                             const int32_t* record_len, int32_t ego_only, uint64_t seed, uint32_t stream_id, float p, void* stream) {
  HMVIT_CHECK_ARG(out && record_len, "dropout: null pointer");
  HMVIT_CHECK_ARG(B > 0 && L > 0 && N > 0 && B * L <= 65535 && N % 4 == 0, "dropout: bad shape (N must be a multiple of 4)");
  HMVIT_CHECK_ARG(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1)");
  DropoutParams q;
  q.a = a; q.resid = resid; q.out = out; q.L = L; q.N = N; q.record_len = record_len; q.ego_only = ego_only ? 1 : 0;
  q.seed_lo = static_cast<uint32_t>(seed); q.seed_hi = static_cast<uint32_t>(seed >> 32); q.stream = stream_id;
  const double th = static_cast<double>(p) * 4294967296.0;
  q.threshold = th >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(th);
  q.scale = 1.0f / (1.0f - p);
  dim3 grid((N / 4 + 255) / 256, 256, B * L);
  dropout_cm_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(q);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// rows: the bf16 [n_rows][256] m-side operand the tensor map describes (M_ROWS) -- any valid pointer otherwise
template <bool M_ROWS>
static int launch_wgrad_tc(const void* rows, long long n_rows, const WgradTcParams& q, cudaStream_t st) {
  using Cfg = WgTc<M_ROWS>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_tc_kernel<M_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (M_ROWS) { int rc = make_weight_tmap(&map, rows, n_rows, 2, 64); if (rc) return rc; }
  wgrad_tc_kernel<M_ROWS><<<num_sms(), Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(map, q);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_bwd_wgrad(const HmvitWgradArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "bwd_wgrad: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->N > 0 && a->B * a->L <= 65535, "bwd_wgrad: bad shape");
  HMVIT_CHECK_ARG(a->N % 32 == 0, "bwd_wgrad: N must be a multiple of 32");
  HMVIT_CHECK_ARG(a->mode && a->record_len && a->a && a->b && a->dw, "bwd_wgrad: null pointer");
  HMVIT_CHECK_ARG(a->dw_rows >= 256 && a->dw_row0 >= 0 && a->dw_row0 + 256 <= a->dw_rows, "bwd_wgrad: bad dw window");
  HMVIT_CHECK_ARG(!(a->b_stats != nullptr && a->b_rows_bf16), "bwd_wgrad: b_stats needs a cm B operand");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tcgen05 form (wgrad_tc.cuh): at least one cm operand, whole 64-token tiles, agent list that fits its shared-memory table
  if (!(a->a_rows_bf16 && a->b_rows_bf16) && a->N % 64 == 0 && a->B * a->L <= 2048) {
    WgradTcParams q;
    q.L = a->L; q.N = a->N; q.n_agents = a->B * a->L; q.mode = a->mode; q.record_len = a->record_len;
    q.ego_only = a->ego_only ? 1 : 0;
    q.m_cm = static_cast<const float*>(a->a); q.n_cm = static_cast<const float*>(a->b);
    q.n_stats = reinterpret_cast<const float2*>(a->b_stats);
    q.dw = a->dw; q.dw_rows = a->dw_rows; q.dw_row0 = a->dw_row0; q.swapped = 0;
    if (a->b_rows_bf16) {        // B rows, A cm: exchange the operands (dW^T = B^T A), the flush writes transposed
      q.n_cm = static_cast<const float*>(a->a); q.m_cm = nullptr; q.swapped = 1;
      return launch_wgrad_tc<true>(a->b, static_cast<long long>(a->B) * a->L * a->N, q, st);
    }
    if (a->a_rows_bf16) return launch_wgrad_tc<true>(a->a, static_cast<long long>(a->B) * a->L * a->N, q, st);
    return launch_wgrad_tc<false>(a->b, 256, q, st);
  }
  WgradParams p;
  p.L = a->L; p.N = a->N; p.mode = a->mode; p.record_len = a->record_len; p.ego_only = a->ego_only ? 1 : 0;
  p.a_cm = static_cast<const float*>(a->a); p.a_rows = static_cast<const __nv_bfloat16*>(a->a);
  p.b_cm = static_cast<const float*>(a->b); p.b_rows = static_cast<const __nv_bfloat16*>(a->b);
  p.b_stats = reinterpret_cast<const float2*>(a->b_stats);
  p.dw = a->dw; p.dw_rows = a->dw_rows; p.dw_row0 = a->dw_row0; p.tok_chunk = 2048;
  dim3 grid(4, (a->N + p.tok_chunk - 1) / p.tok_chunk, a->B * a->L);
  if (a->a_rows_bf16 && a->b_rows_bf16) wgrad_kernel<true, true><<<grid, 256, 0, st>>>(p);
  else if (a->a_rows_bf16) wgrad_kernel<true, false><<<grid, 256, 0, st>>>(p);
  else if (a->b_rows_bf16) wgrad_kernel<false, true><<<grid, 256, 0, st>>>(p);
  else wgrad_kernel<false, false><<<grid, 256, 0, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_bwd_cast_colsum(const float* src, void* dst, float* db, int32_t db_stride, int32_t B, int32_t L, int32_t N,
                                     const int32_t* mode, void* stream) {
  HMVIT_CHECK_ARG(src && dst && db && mode, "bwd_cast_colsum: null pointer");
  HMVIT_CHECK_ARG(B > 0 && L > 0 && N > 0 && db_stride >= 5 * 256, "bwd_cast_colsum: bad shape (db rows hold 5 x 256 sums)");
  const long long R = static_cast<long long>(B) * L * N;
  HMVIT_CHECK_ARG(R < (1ll << 31), "bwd_cast_colsum: too many rows");
  const int rows_per_block = 256;
  dim3 grid(static_cast<unsigned>((R + rows_per_block - 1) / rows_per_block), 5);
  cast_colsum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), db,
                                                                        db_stride, static_cast<int>(R), N, rows_per_block, mode);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_bwd_dgrad_cat(const void* dcat, const void* w0, const void* w1, float* out, int32_t B, int32_t L, int32_t N,
                                   const int32_t* mode, const int32_t* record_len, void* stream) {
  HMVIT_CHECK_ARG(B > 0 && L > 0 && N > 0 && B * L <= DgradCatCfg::MAX_AGENTS, "bwd_dgrad_cat: bad shape (B*L <= 2048)");
  HMVIT_CHECK_ARG(dcat && w0 && w1 && out && mode && record_len, "bwd_dgrad_cat: null pointer");
  const long long R = static_cast<long long>(B) * L * N;
  HMVIT_CHECK_ARG(5 * R + 128 < (1ll << 31), "bwd_dgrad_cat: too many rows for the tensor-map coordinates");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(dgrad_cat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DgradCatCfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  CUtensorMap am, m0, m1;
  int rc = make_weight_tmap(&am, dcat, 5 * R, 2, 128); if (rc) return rc;
  rc = make_weight_tmap(&m0, w0, 1280, 2, 256); if (rc) return rc;
  rc = make_weight_tmap(&m1, w1, 1280, 2, 256); if (rc) return rc;
  DgradCatParams q;
  q.L = L; q.N = N; q.n_agents = B * L; q.R = R; q.mode = mode; q.record_len = record_len; q.out = out;
  dgrad_cat_kernel<<<num_sms(), DgradCatCfg::THREADS, DgradCatCfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(am, m0, m1, q);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

extern "C" int hmvit_group_attn_bwd(const HmvitAttnBwdArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "group_attn_bwd: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->L > 0 && a->H > 0 && a->W > 0, "group_attn_bwd: bad shape");
  HMVIT_CHECK_ARG(a->H % 8 == 0 && a->W % 8 == 0, "group_attn_bwd: H and W must be divisible by the window size 8");
  HMVIT_CHECK_ARG(a->B * a->L <= 65535, "group_attn_bwd: B*L exceeds grid limit");
  HMVIT_CHECK_ARG(a->kind == 0 || a->kind == 1, "group_attn_bwd: kind must be 0 (window) or 1 (grid)");
  HMVIT_CHECK_ARG(a->mode && a->record_len && a->cav_mask && a->T && a->q && a->k && a->v && a->bk && a->bv && a->bias_table &&
                  a->o && a->d_o && a->lse && a->dq && a->dk && a->dv && a->dbk && a->dbv && a->dbias_table,
                  "group_attn_bwd: null pointer");
  HMVIT_CHECK_ARG(a->cell > 0.0, "group_attn_bwd: cell size must be positive");
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(group_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdCfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  AttnBwdParams p;
  p.B = a->B; p.L = a->L; p.H = a->H; p.W = a->W; p.kind = a->kind; p.ego_only = a->ego_only ? 1 : 0;
  p.mode = a->mode; p.record_len = a->record_len; p.cav_mask = a->cav_mask; p.T = a->T; p.cell = a->cell;
  p.q = static_cast<const __nv_bfloat16*>(a->q); p.k = static_cast<const __nv_bfloat16*>(a->k);
  p.v = static_cast<const __nv_bfloat16*>(a->v); p.bk = a->bk; p.bv = a->bv; p.bias_table = a->bias_table;
  p.o = static_cast<const __nv_bfloat16*>(a->o); p.d_o = static_cast<const __nv_bfloat16*>(a->d_o); p.lse = a->lse;
  p.dq = a->dq; p.dk = a->dk; p.dv = a->dv; p.dbk = a->dbk; p.dbv = a->dbv; p.dbias_table = a->dbias_table;
  dim3 grid((a->H / 8) * (a->W / 8) * 2, a->B * a->L);
  group_attn_bwd_kernel<<<grid, AttnBwdCfg::THREADS, AttnBwdCfg::SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// detection decoder (csrc/decoder.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" size_t hmvit_decoder_workspace_bytes(int32_t B, int32_t H, int32_t W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return 2 * align_up(static_cast<size_t>(B) * H * W * 256 * sizeof(__half), 1024);      // two fp16 pixel-row maps (ping-pong)
}
static int make_act_tmap(CUtensorMap* map, const void* act, int B, int H, int W) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) return fail(HMVIT_ERR_CUDA, "hmvit: cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[4] = {256, static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {512, static_cast<cuuint64_t>(W) * 512, static_cast<cuuint64_t>(H) * W * 512};
  cuuint32_t box[4] = {64, DecCfg::TW, DecCfg::BOX_H, 1};   // one half tile per load
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(act), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);      // out-of-range pixels read as zeros: the convolution's padding
  if (r != CUDA_SUCCESS) return fail(HMVIT_ERR_CUDA, "hmvit: cuTensorMapEncodeTiled (activations) failed (" + std::to_string(int(r)) + ")");
  return HMVIT_OK;
}
extern "C" int hmvit_decoder_forward(const HmvitDecoderArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "decoder_forward: null args");
  HMVIT_CHECK_ARG(a->B > 0 && a->H > 0 && a->W > 0 && a->B <= 65535, "decoder_forward: bad shape");
  HMVIT_CHECK_ARG(a->H % 8 == 0 && a->W % 8 == 0, "decoder_forward: H and W must be divisible by 8");
  HMVIT_CHECK_ARG(a->num_convs >= 1 && a->num_convs <= 16, "decoder_forward: num_convs out of range");
  HMVIT_CHECK_ARG(a->anchor_number >= 1 && 8 * a->anchor_number <= kDecMaxOut, "decoder_forward: anchor_number must be in 1..4");
  HMVIT_CHECK_ARG(a->ego_mode && a->x && a->conv_w && a->conv_b && a->head_w && a->head_b && a->psm && a->rm, "decoder_forward: null pointer");
  HMVIT_CHECK_ARG(a->workspace != nullptr && a->workspace_bytes >= hmvit_decoder_workspace_bytes(a->B, a->H, a->W), "decoder_forward: workspace too small");
  HMVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 1023) == 0 && (reinterpret_cast<uintptr_t>(a->conv_w) & 127) == 0,
                  "decoder_forward: workspace must be 1024-byte, conv_w 128-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv3x3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg::SMEM_BYTES);
    if (attr_err == cudaSuccess)
      attr_err = cudaFuncSetAttribute(conv3x3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DecCfg::SMEM_BYTES);
  });
  HMVIT_CHECK_CUDA(attr_err);
  const int N = a->H * a->W;
  const size_t map_bytes = align_up(static_cast<size_t>(a->B) * N * 256 * sizeof(__half), 1024);
  __half* act[2] = {static_cast<__half*>(a->workspace), reinterpret_cast<__half*>(static_cast<uint8_t*>(a->workspace) + map_bytes)};
  nchw_to_nhwc_f16_kernel<<<dim3((N + 63) / 64, a->B), 256, 0, st>>>(a->x, act[0], N);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  const int tiles = ((a->W + DecCfg::TW - 1) / DecCfg::TW) * ((a->H + DecCfg::TH - 1) / DecCfg::TH);
  for (int l = 0; l < a->num_convs; ++l) {
    CUtensorMap mx, mw;
    int rc = make_act_tmap(&mx, act[l & 1], a->B, a->H, a->W); if (rc) return rc;
    rc = make_weight_tmap(&mw, static_cast<const __half*>(a->conv_w) + static_cast<size_t>(l) * 2 * 9 * 256 * 256, 2LL * 9 * 256, 2, 256, true);
    if (rc) return rc;
    ConvParams cp;
    cp.B = a->B; cp.H = a->H; cp.W = a->W; cp.ego_mode = a->ego_mode; cp.bias = a->conv_b + static_cast<size_t>(l) * 2 * 256;
    cp.out = act[(l + 1) & 1];
    cp.head_w = a->head_w; cp.head_b = a->head_b; cp.psm = a->psm; cp.rm = a->rm; cp.n_cls = a->anchor_number;
    if (l == a->num_convs - 1 && a->anchor_number == 2) {
      // the shipped anchor_number: the two 1x1 heads run in the last layer's epilogue (no fp16 round trip of its output)
      conv3x3_kernel<16><<<dim3(tiles, a->B), DecCfg::THREADS, DecCfg::SMEM_BYTES, st>>>(mx, mw, cp);
      HMVIT_CHECK_CUDA(cudaGetLastError());
      return HMVIT_OK;
    }
    conv3x3_kernel<0><<<dim3(tiles, a->B), DecCfg::THREADS, DecCfg::SMEM_BYTES, st>>>(mx, mw, cp);
    HMVIT_CHECK_CUDA(cudaGetLastError());
  }
  HeadsParams hp;
  hp.B = a->B; hp.N = N; hp.n_cls = a->anchor_number; hp.n_reg = 7 * a->anchor_number; hp.ego_mode = a->ego_mode;
  hp.x = act[a->num_convs & 1]; hp.w = a->head_w; hp.bias = a->head_b; hp.psm = a->psm; hp.rm = a->rm;
  det_heads_kernel<<<dim3((N + 127) / 128, a->B), 128, 0, st>>>(hp);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// detection post-processing (csrc/postproc.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" size_t hmvit_postprocess_workspace_bytes(int32_t H, int32_t W, int32_t A) {
  if (H <= 0 || W <= 0 || A <= 0) return 0;
  const size_t n = static_cast<size_t>(H) * W * A;
  return align_up(n, 256) + align_up(n * 4, 256) + align_up(n * 24 * 4, 256) + align_up(kPostMaxCand * 4, 256) +
         align_up(kPostTop * 4, 256) + align_up(static_cast<size_t>(kPostTop) * kPostMaskWords * 8, 256) + 256;
}
extern "C" int hmvit_postprocess(const HmvitPostArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "postprocess: null args");
  HMVIT_CHECK_ARG(a->H > 0 && a->W > 0 && a->A > 0, "postprocess: bad shape");
  HMVIT_CHECK_ARG(a->psm && a->rm && a->anchor_box && a->out_boxes && a->out_scores && a->out_count && a->status, "postprocess: null pointer");
  HMVIT_CHECK_ARG(a->workspace != nullptr && a->workspace_bytes >= hmvit_postprocess_workspace_bytes(a->H, a->W, a->A), "postprocess: workspace too small");
  HMVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "postprocess: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(post_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPostMaxCand * 8);
  });
  HMVIT_CHECK_CUDA(attr_err);
  const size_t n = static_cast<size_t>(a->H) * a->W * a->A;
  PostParams p;
  p.H = a->H; p.W = a->W; p.A = a->A; p.psm = a->psm; p.rm = a->rm; p.anchors = a->anchor_box; p.tmat = a->transformation_matrix;
  p.order_hwl = a->order_hwl ? 1 : 0; p.score_thr = a->score_threshold; p.nms_thr = a->nms_thresh;
  for (int k = 0; k < 4; ++k) p.range[k] = a->range[k];
  uint8_t* ws = static_cast<uint8_t*>(a->workspace);
  p.keep = ws; ws += align_up(n, 256);
  p.score = reinterpret_cast<float*>(ws); ws += align_up(n * 4, 256);
  p.corners = reinterpret_cast<float*>(ws); ws += align_up(n * 24 * 4, 256);
  p.cand = reinterpret_cast<int*>(ws); ws += align_up(kPostMaxCand * 4, 256);
  p.order = reinterpret_cast<int*>(ws); ws += align_up(kPostTop * 4, 256);
  p.mask = reinterpret_cast<unsigned long long*>(ws); ws += align_up(static_cast<size_t>(kPostTop) * kPostMaskWords * 8, 256);
  p.n_cand = reinterpret_cast<int*>(ws); p.n_top = p.n_cand + 1;
  p.out_boxes = a->out_boxes; p.out_scores = a->out_scores; p.out_count = a->out_count; p.status = a->status;
  post_decode_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  post_compact_kernel<<<1, 1024, 0, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  post_sort_kernel<<<1, 1024, kPostMaxCand * 8, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  post_iou_kernel<<<dim3(kPostMaskWords, kPostTop), 64, 0, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  post_nms_kernel<<<1, 32, 0, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

// ------------------------------------------------------------------------------------------------
// PointPillar front end (csrc/pillar.cuh)
// ------------------------------------------------------------------------------------------------
extern "C" int hmvit_pillar_scatter(const HmvitPillarArgs* a, void* stream) {
  HMVIT_CHECK_ARG(a != nullptr, "pillar_scatter: null args");
  HMVIT_CHECK_ARG(a->M >= 0 && a->P >= 1 && a->P <= kPfnMaxPts, "pillar_scatter: 1 <= P <= 32 point slots per pillar");
  HMVIT_CHECK_ARG(a->nx > 0 && a->ny > 0 && a->n_agents > 0, "pillar_scatter: bad canvas shape");
  HMVIT_CHECK_ARG(a->canvas && a->w && a->b, "pillar_scatter: null pointer");
  HMVIT_CHECK_ARG(a->M == 0 || (a->voxel_features && a->voxel_coords && a->voxel_num_points), "pillar_scatter: null pointer");
  HMVIT_CHECK_ARG((reinterpret_cast<uintptr_t>(a->voxel_features) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->voxel_coords) & 15) == 0,
                  "pillar_scatter: voxel_features / voxel_coords must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  HMVIT_CHECK_CUDA(cudaMemsetAsync(a->canvas, 0, static_cast<size_t>(a->n_agents) * kPfnOut * a->ny * a->nx * sizeof(float), st));
  if (a->M == 0) return HMVIT_OK;
  PillarParams p;
  p.M = a->M; p.P = a->P; p.pts = a->voxel_features; p.coords = a->voxel_coords; p.npts = a->voxel_num_points;
  p.w = a->w; p.b = a->b;
  p.vx = a->voxel_size[0]; p.vy = a->voxel_size[1]; p.vz = a->voxel_size[2];
  p.ox = a->offset[0]; p.oy = a->offset[1]; p.oz = a->offset[2];
  p.nx = a->nx; p.ny = a->ny; p.n_agents = a->n_agents; p.canvas = a->canvas; p.channels_last = a->channels_last ? 1 : 0;
  pillar_vfe_scatter_kernel<<<(a->M + 7) / 8, 256, 0, st>>>(p);
  HMVIT_CHECK_CUDA(cudaGetLastError());
  return HMVIT_OK;
}

#ifdef HMVIT_TS
// ------------------------------------------------------------------------------------------------
// timeline instrumentation (tools/build_variant.sh -DHMVIT_TS builds only; not part of the product library)
// ------------------------------------------------------------------------------------------------
extern "C" int hmvit_debug_fa_ts(unsigned long long* host_out /* [2][4][1024] */) {
  HMVIT_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_fa_ts, sizeof(unsigned long long) * 2 * 4 * 1024));
  return HMVIT_OK;
}
extern "C" int hmvit_debug_qkv_ts(unsigned long long* host_out /* [3][512] */) {
  HMVIT_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_qkv_ts, sizeof(unsigned long long) * 3 * 512));
  return HMVIT_OK;
}
extern "C" int hmvit_debug_chain_ts(unsigned long long* host_out /* [2][16][16] */) {
  HMVIT_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_chain_ts, sizeof(unsigned long long) * 2 * 16 * 16));
  return HMVIT_OK;
}
#endif
