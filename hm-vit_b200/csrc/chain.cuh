// Fused post-attention chain for one 128-token tile (sm_100a, persistent, warp-specialised):
//
//   x'  = x + O W_a^T + b_a                       (typed output projection + residual, bf16 operands)
//   x'' = x' + W_2 gelu(W_1 LN'_t(x') + b_1) + b_2   (typed pre-norm FFN + residual, tf32 operands)
//
// Replaces HeteroAttention.to_out (hetero_fusion.py:142-152), the residual (:399 / :442) and
// HeteroPreNormResidual(HeteroFeedForward) (base_transformer.py:129-136, 180-192): five reference ops
// (Linear, add, LayerNorm, Linear-GELU-Linear, add), every intermediate kept on chip.
//
// Data flow per tile (TMEM: D1 = columns 0..255, D2 = columns 256..511):
//   P1  O tile (TMA, bf16) x W_a (TMA ring)            -> D1                       tcgen05 kind::f16
//   E1  D1 + b_a + x (channel-major fp32)              -> x' written back to D1, LN statistics
//   P2  LN'(x') as tf32 32-channel K-chunks (smem ring) x W_1 -> D2                tcgen05 kind::tf32
//   P3  gelu(D2 + b_1) as tf32 K-chunks (same ring) x W_2     -> accumulated ONTO x' in D1
//   E2  D1 + b_2                                       -> x'' stored channel-major
// Warp roles: warps 0-3 transform/epilogue (thread == token row == TMEM lane), warp 4 TMA producer,
// warp 5 MMA issuer (+ TMEM allocator).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace hmvit {

struct ChainParams {
  int B, L, N;
  const int* mode;             // [B*L]
  const int* record_len;       // [B]
  int tile_ego_only;           // 1: only slot 0 of every scene
  const float* resid_cm;       // x   [B*L][256][N]
  float* out_cm;               // x'' [B*L][256][N] (may alias resid_cm)
  const float* ba;             // [2][256]
  const float* ln_gamma;       // [2][256] or null (affine folded into W_1 / b_1 on the host)
  const float* ln_beta;        // [2][256] or null
  float ln_eps;
  const float* b1;             // [2][256]
  const float* b2;             // [2][256]
  float2* stats_out;           // optional [B*L][N] (mean, rstd) of every x'' row: LayerNorm statistics for the next stage
};

struct ChainMaps {             // TMA tensor maps
  CUtensorMap o;               // attention output rows bf16 [B*L*N][256], box 64 x 128
  CUtensorMap wa[2];           // bf16 [256][256], box 64 x 128
  CUtensorMap w1[2];           // fp32 [256][256], box 32 x 128
  CUtensorMap w2[2];
};

struct ChainCfg {
  static constexpr int BM = 128;
  static constexpr int CHUNK = 16384;                 // 128 rows x 128 B
  static constexpr int NF = 4;                        // tf32 A-chunk ring stages
  static constexpr int NS = 4;                        // weight ring stages
  static constexpr int AO_BYTES = 4 * CHUNK;          // O tile, bf16
  static constexpr int SMEM_BYTES = AO_BYTES + NF * CHUNK + NS * CHUNK + 256 + 1024;
  static constexpr int THREADS = 192;
};

__global__ void __launch_bounds__(192, 1) chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  using Cfg = ChainCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sO = smem;
  uint8_t* sF = smem + Cfg::AO_BYTES;
  uint8_t* sW = sF + Cfg::NF * Cfg::CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + Cfg::NS * Cfg::CHUNK);
  uint64_t* w_full = bars;                      // [NS]
  uint64_t* w_empty = w_full + Cfg::NS;         // [NS]
  uint64_t* f_full = w_empty + Cfg::NS;         // [NF]
  uint64_t* f_empty = f_full + Cfg::NF;         // [NF]
  uint64_t* o_full = f_empty + Cfg::NF;
  uint64_t* o_empty = o_full + 1;
  uint64_t* d1_full = o_empty + 1;              // P1 done
  uint64_t* d2_full = d1_full + 1;              // P2 done
  uint64_t* d1_final = d2_full + 1;             // P3 done
  uint64_t* d1_free = d1_final + 1;             // E2 finished reading D1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d1_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NS; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < Cfg::NF; ++s) { mbar_init(&f_full[s], 128); mbar_init(&f_empty[s], 1); }
    mbar_init(o_full, 1); mbar_init(o_empty, 1);
    mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d1_final, 1); mbar_init(d1_free, 128);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t D1 = tmem_base, D2 = tmem_base + 256;

  const int tiles_per_agent = (p.N + Cfg::BM - 1) / Cfg::BM;
  const int total_tiles = p.B * p.L * tiles_per_agent;

  auto tile_agent = [&](int t, int& a, int& tok0) -> bool {
    a = t / tiles_per_agent;
    tok0 = (t - a * tiles_per_agent) * Cfg::BM;
    const int b = a / p.L, l = a - b * p.L;
    return l < p.record_len[b] && !(p.tile_ego_only && l != 0);
  };

  if (warp < 4) {
    // ============================ transform / epilogue warps ============================
    const int row = threadIdx.x;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t ti = 0, itf = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int a, tok0;
      if (!tile_agent(t, a, tok0)) continue;
      const int type = p.mode[a] != 0 ? 1 : 0;
      const int tok = tok0 + row;
      const bool valid = tok < p.N;
      const size_t cm_off = static_cast<size_t>(a) * kC * p.N + (valid ? tok : 0);
      const float* res = p.resid_cm + cm_off;
      float* dst = p.out_cm + cm_off;
      // ---- E1: x' = D1 + b_a + x ; statistics.  The residual loads run 3 pieces ahead and start
      //      before the projection MMAs have finished. ----
      float rv[3][32];
#pragma unroll
      for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int k = 0; k < 32; ++k) rv[q][k] = valid ? res[static_cast<size_t>(q * 32 + k) * p.N] : 0.f;
      }
      mbar_wait(d1_full, ti & 1);
      tc_fence_after();
      float s0 = 0.f, sum = 0.f, sq = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint32_t r[32];
        tmem_ld32(D1 + lane_base + q * 32, r);
        tmem_ld_wait();
        const float* bias = p.ba + type * kC + q * 32;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float v = __uint_as_float(r[k]) + __ldg(bias + k) + rv[q % 3][k];
          if (q == 0 && k == 0) s0 = v;
          const float d = v - s0;
          sum += d; sq += d * d;
          r[k] = __float_as_uint(v);
        }
        tmem_st32(D1 + lane_base + q * 32, r);
        if (q + 3 < 8) {
#pragma unroll
          for (int k = 0; k < 32; ++k) rv[q % 3][k] = valid ? res[static_cast<size_t>((q + 3) * 32 + k) * p.N] : 0.f;
        }
      }
      tmem_st_wait();
      const float md = sum * (1.0f / kC);
      const float mean = s0 + md;
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / kC) - md * md, 0.f) + p.ln_eps);
      // ---- P2 feed: LN'(x') -> tf32 K-chunks ----
      const bool affine = p.ln_gamma != nullptr;
      const float* gam = affine ? p.ln_gamma + type * kC : nullptr;
      const float* bet = affine ? p.ln_beta + type * kC : nullptr;
      const float nmr = -mean * rstd;
#pragma unroll 1
      for (int kc = 0; kc < 8; ++kc, ++itf) {
        const uint32_t fs = itf % Cfg::NF, ph = (itf / Cfg::NF) & 1u;
        uint32_t r[32];
        tmem_ld32(D1 + lane_base + kc * 32, r);
        tmem_ld_wait();
        if (affine) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            r[k] = __float_as_uint(tf32_rn(fmaf(__uint_as_float(r[k]), rstd, nmr) * __ldg(gam + kc * 32 + k) + __ldg(bet + kc * 32 + k)));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(tf32_rn(fmaf(__uint_as_float(r[k]), rstd, nmr)));
        }
        mbar_wait(&f_empty[fs], ph ^ 1u);
        uint8_t* dstF = sF + fs * Cfg::CHUNK;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<uint4*>(dstF + sw128_offset(row, u)) = make_uint4(r[u * 4], r[u * 4 + 1], r[u * 4 + 2], r[u * 4 + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&f_full[fs]);
      }
      // ---- P3 feed: gelu(D2 + b_1) -> tf32 K-chunks ----
      mbar_wait(d2_full, ti & 1);
      tc_fence_after();
      const float* b1 = p.b1 + type * kC;
#pragma unroll 1
      for (int kc = 0; kc < 8; ++kc, ++itf) {
        const uint32_t fs = itf % Cfg::NF, ph = (itf / Cfg::NF) & 1u;
        uint32_t r[32];
        tmem_ld32(D2 + lane_base + kc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(tf32_rn(gelu_erf_fast(__uint_as_float(r[k]) + __ldg(b1 + kc * 32 + k))));
        mbar_wait(&f_empty[fs], ph ^ 1u);
        uint8_t* dstF = sF + fs * Cfg::CHUNK;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<uint4*>(dstF + sw128_offset(row, u)) = make_uint4(r[u * 4], r[u * 4 + 1], r[u * 4 + 2], r[u * 4 + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&f_full[fs]);
      }
      // ---- E2: x'' = D1 + b_2 (+ LayerNorm statistics of x'' for the next stage) ----
      mbar_wait(d1_final, ti & 1);
      tc_fence_after();
      const float* b2 = p.b2 + type * kC;
      float t0 = 0.f, tsum = 0.f, tsq = 0.f;
#pragma unroll 1
      for (int q = 0; q < 8; ++q) {
        uint32_t r[32];
        tmem_ld32(D1 + lane_base + q * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float v = __uint_as_float(r[k]) + __ldg(b2 + q * 32 + k);
          if (q == 0 && k == 0) t0 = v;
          const float d = v - t0;
          tsum += d; tsq += d * d;
          if (valid) dst[static_cast<size_t>(q * 32 + k) * p.N] = v;
        }
      }
      if (p.stats_out != nullptr && valid) {
        const float tmd = tsum * (1.0f / kC);
        p.stats_out[static_cast<size_t>(a) * p.N + tok] =
            make_float2(t0 + tmd, rsqrtf(fmaxf(tsq * (1.0f / kC) - tmd * tmd, 0.f) + p.ln_eps));
      }
      tc_fence_before();
      mbar_arrive(d1_free);
      ++ti;
    }
  } else if (warp == 4) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      tma_prefetch_desc(&maps.o);
      uint32_t ti = 0, itw = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int a, tok0;
        if (!tile_agent(t, a, tok0)) continue;
        const int type = p.mode[a] != 0 ? 1 : 0;
        mbar_wait(o_empty, (ti & 1) ^ 1u);
        mbar_arrive_expect_tx(o_full, Cfg::AO_BYTES);
        const int row0 = a * p.N + tok0;
        for (int kc = 0; kc < 4; ++kc) tma_load_2d(sO + kc * Cfg::CHUNK, &maps.o, o_full, kc * 64, row0);
        auto wstage = [&](const CUtensorMap* m, int c0, int c1) {
          const uint32_t s = itw % Cfg::NS, ph = (itw / Cfg::NS) & 1u;
          mbar_wait(&w_empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&w_full[s], Cfg::CHUNK);
          tma_load_2d(sW + s * Cfg::CHUNK, m, &w_full[s], c0, c1);
          ++itw;
        };
        for (int nc = 0; nc < 2; ++nc)
          for (int kc = 0; kc < 4; ++kc) wstage(&maps.wa[type], kc * 64, nc * 128);
        for (int kc = 0; kc < 8; ++kc)
          for (int nc = 0; nc < 2; ++nc) wstage(&maps.w1[type], kc * 32, nc * 128);
        for (int kc = 0; kc < 8; ++kc)
          for (int nc = 0; nc < 2; ++nc) wstage(&maps.w2[type], kc * 32, nc * 128);
        ++ti;
      }
    }
  } else {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_bf16 = umma_idesc(1u, 128, 128);
      constexpr uint32_t idesc_tf32 = umma_idesc(2u, 128, 128);
      const uint32_t o_base = smem_u32(sO), f_base = smem_u32(sF), w_base = smem_u32(sW);
      uint32_t ti = 0, itw = 0, itf = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int a, tok0;
        if (!tile_agent(t, a, tok0)) continue;
        mbar_wait(d1_free, (ti & 1) ^ 1u);
        mbar_wait(o_full, ti & 1);
        tc_fence_after();
        // P1: D1 = O W_a^T
        for (int nc = 0; nc < 2; ++nc) {
          for (int kc = 0; kc < 4; ++kc, ++itw) {
            const uint32_t s = itw % Cfg::NS, ph = (itw / Cfg::NS) & 1u;
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss<2>(D1 + nc * 128, umma_desc_sw128(o_base + kc * Cfg::CHUNK + ks * 32),
                         umma_desc_sw128(w_base + s * Cfg::CHUNK + ks * 32), idesc_bf16, (kc | ks) != 0 ? 1u : 0u);
            umma_commit(&w_empty[s]);
          }
        }
        umma_commit(o_empty);
        umma_commit(d1_full);
        // P2: D2 = LN'(x') W_1^T      P3: D1 += gelu(.) W_2^T
        for (int phase = 0; phase < 2; ++phase) {
          const uint32_t dacc = phase == 0 ? D2 : D1;
          for (int kc = 0; kc < 8; ++kc, ++itf) {
            const uint32_t fs = itf % Cfg::NF, fph = (itf / Cfg::NF) & 1u;
            mbar_wait(&f_full[fs], fph);
            tc_fence_after();
            for (int nc = 0; nc < 2; ++nc, ++itw) {
              const uint32_t s = itw % Cfg::NS, ph = (itw / Cfg::NS) & 1u;
              mbar_wait(&w_full[s], ph);
              tc_fence_after();
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_ss<4>(dacc + nc * 128, umma_desc_sw128(f_base + fs * Cfg::CHUNK + ks * 32),
                           umma_desc_sw128(w_base + s * Cfg::CHUNK + ks * 32), idesc_tf32,
                           (phase == 1 || (kc | ks) != 0) ? 1u : 0u);
              umma_commit(&w_empty[s]);
            }
            umma_commit(&f_empty[fs]);
          }
          umma_commit(phase == 0 ? d2_full : d1_final);
        }
        ++ti;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace hmvit
