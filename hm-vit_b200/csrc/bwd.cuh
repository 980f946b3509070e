// Backward-pass kernels of the HM-ViT fusion path that are not GEMM-shaped dgrads (those reuse the
// tcgen05 row-GEMM of rowgemm.cuh with transposed weights): LayerNorm statistics / backward,
// GELU backward, casts, bias-gradient column sums and the typed weight-gradient GEMM.
//
// The reference has no hand-written backward: autograd differentiates
//   HeteroLayerNorm / HeteroFeedForward / HeteroPreNormResidual   base_transformer.py:129-192
//   HeteroAttention.to_qkv / to_out                               hetero_fusion.py:111-152
// These kernels are the adjoint of the restructured forward: gradients are produced for
// the FOLDED weights (W_cat, W_a, W_1', W_2, ...) and pulled back to the module parameters on the
// host through the (tiny, differentiable) folding function.
//
// Layouts as in the forward: "cm" = fp32 [agents][256][N] (the module's (B, L, C, H, W)), "rows" =
// bf16 [agents*N][256].  All kernels are typed by mode[a] and skip padded agent slots.
#pragma once
#include "common.cuh"
#include <mma.h>

namespace hmvit {

HMVIT_DEVINL bool agent_active(int a, int L, const int* __restrict__ record_len, int ego_only) {
  const int b = a / L, l = a - b * L;
  return l < record_len[b] && !(ego_only && l != 0);
}

// ------------------------------------------------------------------------------------------
// per-token LayerNorm statistics of a cm tensor: stats[a*N + tok] = (mean, rstd) over the 256 channels
// (biased variance, like nn.LayerNorm; shifted single pass like the forward kernels)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) row_stats_kernel(const float* __restrict__ x, float2* __restrict__ stats, int L, int N,
                                                        const int* __restrict__ record_len, int ego_only, float eps) {
  const int a = blockIdx.y;
  if (!agent_active(a, L, record_len, ego_only)) return;
  const int tok = blockIdx.x * 128 + threadIdx.x;
  if (tok >= N) return;
  const float* src = x + static_cast<size_t>(a) * kC * N + tok;
  const float s0 = __ldg(src);
  float sum = 0.f, sq = 0.f;
#pragma unroll 8
  for (int c = 0; c < kC; ++c) { const float d = __ldg(src + static_cast<size_t>(c) * N) - s0; sum += d; sq += d * d; }
  const float md = sum * (1.0f / kC);
  stats[static_cast<size_t>(a) * N + tok] = make_float2(s0 + md, rsqrtf(fmaxf(sq * (1.0f / kC) - md * md, 0.f) + eps));
}

// ------------------------------------------------------------------------------------------
// LayerNorm backward (no affine: gamma / beta are folded into the consuming GEMM) + residual:
//   z = (x - mean) rstd ;  dx = dres + rstd (dz - mean_c(dz) - z mean_c(dz z))
// dx may alias dres.  thread == token, channel loop (every access is a coalesced 512-byte warp row).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ln_bwd_cm_kernel(const float* __restrict__ dz, const float* __restrict__ x,
                                                        const float2* __restrict__ stats, const float* dres, float* dx,
                                                        int L, int N, const int* __restrict__ record_len, int ego_only) {
  const int a = blockIdx.y;
  if (!agent_active(a, L, record_len, ego_only)) return;
  const int tok = blockIdx.x * 128 + threadIdx.x;
  if (tok >= N) return;
  const size_t base = static_cast<size_t>(a) * kC * N + tok;
  const float2 st = stats[static_cast<size_t>(a) * N + tok];
  const float mean = st.x, rstd = st.y;
  float m1 = 0.f, m2 = 0.f;
#pragma unroll 8
  for (int c = 0; c < kC; ++c) {
    const float g = __ldg(dz + base + static_cast<size_t>(c) * N);
    const float z = (__ldg(x + base + static_cast<size_t>(c) * N) - mean) * rstd;
    m1 += g; m2 = fmaf(g, z, m2);
  }
  m1 *= (1.0f / kC); m2 *= (1.0f / kC);
#pragma unroll 8
  for (int c = 0; c < kC; ++c) {
    const size_t o = base + static_cast<size_t>(c) * N;
    const float g = __ldg(dz + o);
    const float z = (__ldg(x + o) - mean) * rstd;
    dx[o] = dres[o] + rstd * (g - m1 - z * m2);
  }
}

// ------------------------------------------------------------------------------------------
// GELU (erf form) backward, in place: hp <- gelu(hp) (operand of the W_2 weight gradient),
// dh <- dh * gelu'(hp) (gradient w.r.t. the pre-activation)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gelu_bwd_kernel(float4* __restrict__ hp, float4* __restrict__ dh, size_t n4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 h = hp[i], d = dh[i];
    float* hv = reinterpret_cast<float*>(&h);
    float* dv = reinterpret_cast<float*>(&d);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = hv[e];
      const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * __expf(-0.5f * v * v);
      hv[e] = v * cdf;
      dv[e] = dv[e] * fmaf(v, pdf, cdf);
    }
    hp[i] = h; dh[i] = d;
  }
}

// fp32 -> bf16 (round to nearest even), flat
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, size_t n4) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    dst[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ------------------------------------------------------------------------------------------
// bias gradients: db[type][c] += sum over the tokens of every active agent of that type
// ------------------------------------------------------------------------------------------
// cm source: one warp per channel row of N contiguous floats
__global__ void __launch_bounds__(256) colsum_cm_kernel(const float* __restrict__ y, float* __restrict__ db, int db_stride,
                                                        int L, int N, const int* __restrict__ mode,
                                                        const int* __restrict__ record_len, int ego_only) {
  const int a = blockIdx.y;
  if (!agent_active(a, L, record_len, ego_only)) return;
  const int type = mode[a] != 0 ? 1 : 0;
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const float* src = y + (static_cast<size_t>(a) * kC + c) * N;
  float s = 0.f;
  for (int t = lane; t < N; t += 32) s += __ldg(src + t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) atomicAdd(db + static_cast<size_t>(type) * db_stride + c, s);
}
// rows source (bf16 [rows][256]): thread == channel, 256 tokens per block
__global__ void __launch_bounds__(256) colsum_rows_kernel(const __nv_bfloat16* __restrict__ y, float* __restrict__ db, int db_stride,
                                                          int L, int N, const int* __restrict__ mode,
                                                          const int* __restrict__ record_len, int ego_only) {
  const int a = blockIdx.y;
  if (!agent_active(a, L, record_len, ego_only)) return;
  const int type = mode[a] != 0 ? 1 : 0;
  const int t0 = blockIdx.x * 256, t1 = min(N, t0 + 256);
  const __nv_bfloat16* src = y + (static_cast<size_t>(a) * N + t0) * kC + threadIdx.x;
  float s = 0.f;
  for (int t = t0; t < t1; ++t, src += kC) s += __bfloat162float(*src);
  atomicAdd(db + static_cast<size_t>(type) * db_stride + threadIdx.x, s);
}

// ------------------------------------------------------------------------------------------
// typed weight gradient:  dW[type][m][n] += sum_tok A(tok, m) B(tok, n),   m, n in [0, 256)
// over the tokens of every active agent of that type.  Operands come either from a cm tensor
// (fp32 [a][256][N], optionally normalised on the fly with per-token (mean, rstd)) or from bf16 rows.
// Warp-level tensor-core tiles (wmma tf32 m16n16k8, fp32 accumulate): one CTA = a 128 x 128 tile of dW
// over a chunk of tokens of one agent, partial sums reduced with fp32 atomics.
// ------------------------------------------------------------------------------------------
struct WgradParams {
  int L, N;
  const int* mode;
  const int* record_len;
  int ego_only;
  const float* a_cm;              // A operand, cm (A_ROWS == false)
  const __nv_bfloat16* a_rows;    // A operand, bf16 rows (A_ROWS == true)
  const float* b_cm;              // B operand, cm
  const __nv_bfloat16* b_rows;    // B operand, bf16 rows
  const float2* b_stats;          // optional (B cm only): normalise B with per-token (mean, rstd)
  float* dw;                      // [2][dw_rows][256] fp32, accumulated
  int dw_rows;                    // rows of one type's matrix (256 * planes)
  int dw_row0;                    // first row this call accumulates into
  int tok_chunk;                  // tokens per CTA (multiple of 32)
};

constexpr int kWgKT = 32;                       // tokens per K step
constexpr int kWgLdK = kWgKT + 4;               // [128][36]  (cm source: k contiguous)
constexpr int kWgLdM = 128 + 4;                 // [32][132]  (rows source: m contiguous)
constexpr int kWgTile = 128 * kWgLdK;           // floats per operand tile (>= 32 * 132)

template <bool A_ROWS, bool B_ROWS>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradParams p) {
  using namespace nvcuda;
  const int a = blockIdx.z;
  if (!agent_active(a, p.L, p.record_len, p.ego_only)) return;
  const int type = p.mode[a] != 0 ? 1 : 0;
  const int m0 = (blockIdx.x >> 1) * 128, n0 = (blockIdx.x & 1) * 128;
  const int tok_begin = blockIdx.y * p.tok_chunk;
  const int tok_end = min(p.N, tok_begin + p.tok_chunk);
  if (tok_begin >= tok_end) return;

  __shared__ __align__(128) float sA[kWgTile];
  __shared__ __align__(128) float sB[kWgTile];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;      // warp tile: 64 (m) x 32 (n)

  wmma::fragment<wmma::accumulator, 16, 16, 8, float> acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::fill_fragment(acc[i][j], 0.f);

  for (int tok0 = tok_begin; tok0 < tok_end; tok0 += kWgKT) {
    // ---- stage A: 128 (m) x 32 (tok) ----
    if constexpr (A_ROWS) {
      // rows [tok][256]: 32 tokens x 16 uint4 (8 channels each) -> sA[k][m]
      for (int e = threadIdx.x; e < 32 * 16; e += 256) {
        const int k = e >> 4, u = e & 15;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.a_rows + (static_cast<size_t>(a) * p.N + tok0 + k) * kC + m0) + u);
        float* d = sA + k * kWgLdM + u * 8;
        d[0] = bf16_lo(v.x); d[1] = bf16_hi(v.x); d[2] = bf16_lo(v.y); d[3] = bf16_hi(v.y);
        d[4] = bf16_lo(v.z); d[5] = bf16_hi(v.z); d[6] = bf16_lo(v.w); d[7] = bf16_hi(v.w);
      }
    } else {
      // cm [256][N]: 128 channel rows x 8 float4 -> sA[m][k]
      for (int e = threadIdx.x; e < 128 * 8; e += 256) {
        const int m = e >> 3, u = e & 7;
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.a_cm + (static_cast<size_t>(a) * kC + m0 + m) * p.N + tok0) + u);
        *reinterpret_cast<float4*>(sA + m * kWgLdK + u * 4) = v;
      }
    }
    // ---- stage B: 128 (n) x 32 (tok) ----
    if constexpr (B_ROWS) {
      for (int e = threadIdx.x; e < 32 * 16; e += 256) {
        const int k = e >> 4, u = e & 15;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.b_rows + (static_cast<size_t>(a) * p.N + tok0 + k) * kC + n0) + u);
        float* d = sB + k * kWgLdM + u * 8;
        d[0] = bf16_lo(v.x); d[1] = bf16_hi(v.x); d[2] = bf16_lo(v.y); d[3] = bf16_hi(v.y);
        d[4] = bf16_lo(v.z); d[5] = bf16_hi(v.z); d[6] = bf16_lo(v.w); d[7] = bf16_hi(v.w);
      }
    } else {
      for (int e = threadIdx.x; e < 128 * 8; e += 256) {
        const int n = e >> 3, u = e & 7;
        float4 v = __ldg(reinterpret_cast<const float4*>(p.b_cm + (static_cast<size_t>(a) * kC + n0 + n) * p.N + tok0) + u);
        if (p.b_stats != nullptr) {
          const float4 s01 = __ldg(reinterpret_cast<const float4*>(p.b_stats + static_cast<size_t>(a) * p.N + tok0 + u * 4));
          const float4 s23 = __ldg(reinterpret_cast<const float4*>(p.b_stats + static_cast<size_t>(a) * p.N + tok0 + u * 4) + 1);
          v.x = (v.x - s01.x) * s01.y; v.y = (v.y - s01.z) * s01.w;
          v.z = (v.z - s23.x) * s23.y; v.w = (v.w - s23.z) * s23.w;
        }
        *reinterpret_cast<float4*>(sB + n * kWgLdK + u * 4) = v;
      }
    }
    __syncthreads();
    // ---- tensor-core phase: 4 k-steps of 8 tokens ----
#pragma unroll
    for (int kk = 0; kk < kWgKT; kk += 8) {
      wmma::fragment<wmma::matrix_a, 16, 16, 8, wmma::precision::tf32, typename std::conditional<A_ROWS, wmma::col_major, wmma::row_major>::type> fa[4];
      wmma::fragment<wmma::matrix_b, 16, 16, 8, wmma::precision::tf32, typename std::conditional<B_ROWS, wmma::row_major, wmma::col_major>::type> fb[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if constexpr (A_ROWS) wmma::load_matrix_sync(fa[i], sA + kk * kWgLdM + wm + i * 16, kWgLdM);
        else wmma::load_matrix_sync(fa[i], sA + (wm + i * 16) * kWgLdK + kk, kWgLdK);
#pragma unroll
        for (int t = 0; t < fa[i].num_elements; ++t) fa[i].x[t] = wmma::__float_to_tf32(fa[i].x[t]);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if constexpr (B_ROWS) wmma::load_matrix_sync(fb[j], sB + kk * kWgLdM + wn + j * 16, kWgLdM);
        else wmma::load_matrix_sync(fb[j], sB + (wn + j * 16) * kWgLdK + kk, kWgLdK);
#pragma unroll
        for (int t = 0; t < fb[j].num_elements; ++t) fb[j].x[t] = wmma::__float_to_tf32(fb[j].x[t]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) wmma::mma_sync(acc[i][j], fa[i], fb[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: fragment -> per-warp 16 x 16 patch in smem -> coalesced fp32 atomics ----
  float* patch = sA + warp * 256;                  // sA is free after the final __syncthreads
  float* dst = p.dw + (static_cast<size_t>(type) * p.dw_rows + p.dw_row0 + m0 + wm) * kC + n0 + wn;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      wmma::store_matrix_sync(patch, acc[i][j], 16, wmma::mem_row_major);
      __syncwarp();
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int idx = e * 32 + lane, r = idx >> 4, c = idx & 15;
        atomicAdd(dst + static_cast<size_t>(i * 16 + r) * kC + j * 16 + c, patch[idx]);
      }
      __syncwarp();
    }
  }
}

}  // namespace hmvit
