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
#include "wmma_shared.cuh"

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

// fp32 -> bf16 of the five gradient planes [5][R][256] of the fused Q | K' | V' projection AND their typed column sums in
// the same pass (the bias gradient of the projection): db[type(agent of the row)][p * 256 + c] += src[p][row][c].  Replaces
// cast_bf16_kernel + five colsum_rows launches that re-read the bf16 copy.  Rows of padded slots must be zero (they are:
// the attention backward accumulates into a zero-filled buffer and never touches them), so no validity test is needed.
// grid (ceil(R / rows_per_block), 5), 256 threads: thread = (row lane tid / 64, 4-channel group tid % 64).
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, float* __restrict__ db,
                                                          int db_stride, int R, int N, int rows_per_block, const int* __restrict__ mode) {
  const int pl = blockIdx.y, cg = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  const size_t base = static_cast<size_t>(pl) * R;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
  int ag = r0 / N, nb = (ag + 1) * N;                            // agent of the current row, first row of the next agent
  bool t1 = __ldg(mode + ag) != 0;
#pragma unroll 4
  for (int r = r0 + rl; r < r1; r += 4) {
    const float4 v = __ldg(src + (base + r) * 64 + cg);
    dst[(base + r) * 64 + cg] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    while (r >= nb) { ++ag; nb += N; t1 = __ldg(mode + ag) != 0; }
    if (t1) { a1.x += v.x; a1.y += v.y; a1.z += v.z; a1.w += v.w; }
    else { a0.x += v.x; a0.y += v.y; a0.z += v.z; a0.w += v.w; }
  }
  __shared__ float4 sAcc[2][4][64];
  sAcc[0][rl][cg] = a0; sAcc[1][rl][cg] = a1;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int t = threadIdx.x >> 6;
    float4 s = sAcc[t][0][cg];
#pragma unroll
    for (int i = 1; i < 4; ++i) { const float4 o = sAcc[t][i][cg]; s.x += o.x; s.y += o.y; s.z += o.z; s.w += o.w; }
    float* d = db + static_cast<size_t>(t) * db_stride + pl * kC + cg * 4;
    atomicAdd(d, s.x); atomicAdd(d + 1, s.y); atomicAdd(d + 2, s.z); atomicAdd(d + 3, s.w);
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
// Warp-level tensor-core tiles (wmma bf16 m16n16k16, fp32 accumulate): one CTA = a 128 x 128 tile of dW
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

// Operands are converted to bf16 while they are staged (fp32 accumulate): half the shared-memory traffic and
// twice the tensor-core rate of tf32 fragments.  The next K tile is prefetched into registers while the current
// one is multiplied (two shared-memory buffers, one barrier per tile).
constexpr int kWgKT = 32;                       // tokens per K tile
constexpr int kWgLdK = kWgKT + 8;               // [128][40] bf16  (cm source: k contiguous)
constexpr int kWgLdM = 128 + 8;                 // [32][136] bf16  (rows source: m contiguous)
constexpr int kWgTile = 128 * kWgLdK;           // bf16 elements per operand tile (5120 >= 32 * 136 = 4352)

template <bool ROWS>
struct WgStage {                                // register image of one operand tile: 128 x 32 elements
  uint4 v[ROWS ? 2 : 4];                        // rows: 2 x 8 bf16 per thread; cm: 4 x float4 per thread
  float4 st[2];                                 // cm + stats: (mean, rstd) of this thread's 4 tokens (the same for its 4 rows)
};

template <bool ROWS>
HMVIT_DEVINL void wg_load(WgStage<ROWS>& r, const float* cm, const __nv_bfloat16* rows, const float2* stats, int a, int N, int c0,
                          int tok0, int tid) {
  if constexpr (ROWS) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * 256, k = e >> 4, u = e & 15;
      r.v[i] = __ldg(reinterpret_cast<const uint4*>(rows + (static_cast<size_t>(a) * N + tok0 + k) * kC + c0) + u);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256, m = e >> 3, u = e & 7;
      const float4 f = __ldg(reinterpret_cast<const float4*>(cm + (static_cast<size_t>(a) * kC + c0 + m) * N + tok0) + u);
      r.v[i] = make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w));
      if (i == 0 && stats != nullptr) {           // u = tid & 7 for every i: one pair of loads per tile
        const float4* sp = reinterpret_cast<const float4*>(stats + static_cast<size_t>(a) * N + tok0 + u * 4);
        r.st[0] = __ldg(sp); r.st[1] = __ldg(sp + 1);
      }
    }
  }
}

template <bool ROWS>
HMVIT_DEVINL void wg_store(const WgStage<ROWS>& r, __nv_bfloat16* sm, bool has_stats, int tid) {
  if constexpr (ROWS) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int e = tid + i * 256, k = e >> 4, u = e & 15;
      *reinterpret_cast<uint4*>(sm + k * kWgLdM + u * 8) = r.v[i];                 // sm[k][m]
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256, m = e >> 3, u = e & 7;
      float x0 = __uint_as_float(r.v[i].x), x1 = __uint_as_float(r.v[i].y), x2 = __uint_as_float(r.v[i].z), x3 = __uint_as_float(r.v[i].w);
      if (has_stats) {
        const float4 s01 = r.st[0], s23 = r.st[1];
        x0 = (x0 - s01.x) * s01.y; x1 = (x1 - s01.z) * s01.w; x2 = (x2 - s23.x) * s23.y; x3 = (x3 - s23.z) * s23.w;
      }
      *reinterpret_cast<uint2*>(sm + m * kWgLdK + u * 4) = make_uint2(pack_bf16x2(x0, x1), pack_bf16x2(x2, x3));   // sm[m][k]
    }
  }
}

#ifndef HMVIT_WG_MINB
#define HMVIT_WG_MINB 1
#endif
template <bool A_ROWS, bool B_ROWS>
__global__ void __launch_bounds__(256, HMVIT_WG_MINB) wgrad_kernel(const WgradParams p) {
  using namespace nvcuda;
  const int a = blockIdx.z;
  if (!agent_active(a, p.L, p.record_len, p.ego_only)) return;
  const int type = p.mode[a] != 0 ? 1 : 0;
  const int m0 = (blockIdx.x >> 1) * 128, n0 = (blockIdx.x & 1) * 128;
  const int tok_begin = blockIdx.y * p.tok_chunk;
  const int tok_end = min(p.N, tok_begin + p.tok_chunk);
  if (tok_begin >= tok_end) return;

  __shared__ __align__(128) __nv_bfloat16 sA[2][kWgTile];
  __shared__ __align__(128) __nv_bfloat16 sB[2][kWgTile];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;      // warp tile: 64 (m) x 32 (n)
  const bool has_stats = !B_ROWS && p.b_stats != nullptr;

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::fill_fragment(acc[i][j], 0.f);

  WgStage<A_ROWS> ra;
  WgStage<B_ROWS> rb;
  wg_load<A_ROWS>(ra, p.a_cm, p.a_rows, nullptr, a, p.N, m0, tok_begin, tid);
  wg_load<B_ROWS>(rb, p.b_cm, p.b_rows, p.b_stats, a, p.N, n0, tok_begin, tid);
  int buf = 0;
  for (int tok0 = tok_begin; tok0 < tok_end; tok0 += kWgKT, buf ^= 1) {
    wg_store<A_ROWS>(ra, sA[buf], false, tid);
    wg_store<B_ROWS>(rb, sB[buf], has_stats, tid);
    __syncthreads();                       // tile `buf` is complete; every warp is past its reads of tile buf (2 tiles ago)
    if (tok0 + kWgKT < tok_end) {          // prefetch the next tile into registers while this one is multiplied
      wg_load<A_ROWS>(ra, p.a_cm, p.a_rows, nullptr, a, p.N, m0, tok0 + kWgKT, tid);
      wg_load<B_ROWS>(rb, p.b_cm, p.b_rows, p.b_stats, a, p.N, n0, tok0 + kWgKT, tid);
    }
    const __nv_bfloat16* tA = sA[buf];
    const __nv_bfloat16* tB = sB[buf];
#pragma unroll
    for (int kk = 0; kk < kWgKT; kk += 16) {
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, typename std::conditional<B_ROWS, wmma::row_major, wmma::col_major>::type> fb[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if constexpr (B_ROWS) wmma_load_shared(fb[j], tB + kk * kWgLdM + wn + j * 16, kWgLdM);
        else wmma_load_shared(fb[j], tB + (wn + j * 16) * kWgLdK + kk, kWgLdK);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {          // one A fragment live at a time (register budget: 2 CTAs / SM)
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, typename std::conditional<A_ROWS, wmma::col_major, wmma::row_major>::type> fa;
        if constexpr (A_ROWS) wmma_load_shared(fa, tA + kk * kWgLdM + wm + i * 16, kWgLdM);
        else wmma_load_shared(fa, tA + (wm + i * 16) * kWgLdK + kk, kWgLdK);
#pragma unroll
        for (int j = 0; j < 2; ++j) wmma::mma_sync(acc[i][j], fa, fb[j], acc[i][j]);
      }
    }
  }
  __syncthreads();

  // ---- epilogue: fragment -> per-warp 16 x 16 fp32 patch in smem -> coalesced fp32 atomics ----
  float* patch = reinterpret_cast<float*>(&sA[0][0]) + warp * 256;      // 8 KB of the 20 KB of sA
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
