// Fused post-attention chain for one 128-token tile (sm_100a, persistent, warp-specialised):
//
//   x'  = x + O W_a^T + b_a                       (typed output projection + residual, bf16 operands)
//   x'' = x' + W_2 gelu(W_1 LN'_t(x') + b_1) + b_2   (typed pre-norm FFN + residual, fp16 operands: the same 11-bit
//                                                      significand as tf32 at twice the MMA rate and half the operand
//                                                      bytes; conversions saturate at +-65504 -- LN outputs are
//                                                      bounded by 16, measured parity identical to tf32)
//
// Replaces HeteroAttention.to_out (hetero_fusion.py:142-152), the residual (:399 / :442) and
// HeteroPreNormResidual(HeteroFeedForward) (base_transformer.py:129-136, 180-192): five reference ops
// (Linear, add, LayerNorm, Linear-GELU-Linear, add), every intermediate kept on chip.
//
// Data flow per tile (TMEM: D1 = columns 0..255, D2 = columns 256..511):
//   P1  O tile (TMA, bf16) x W_a (TMA ring)            -> D1                       tcgen05 kind::f16
//   E1  D1 + b_a + x (channel-major fp32)              -> x' written back to D1, LN statistics
//   P2  LN'(x') as fp16 64-channel K-chunks (smem ring) x W_1 -> D2                tcgen05 kind::f16
//   P3  gelu(D2 + b_1) as fp16 K-chunks (same ring) x W_2     -> accumulated ONTO x' in D1
//   E2  D1 + b_2                                       -> x'' stored channel-major (+ LN statistics)
// Warp roles (18 warps): warps 0-15 transform / epilogue in 4 groups of 4 warps -- thread == token row
// == TMEM lane (lane quarter = warp % 4), group g owns columns [64g, 64g+64) in E1 / E2 and the
// K-chunk g in the P2 / P3 feeds, so four chunks are produced concurrently and every SM
// sub-partition has four transform warps to hide TMEM / global latency; warp 16 TMA producer;
// warp 17 MMA issuer (+ TMEM allocator).
#pragma once
#include "common.cuh"
#include <cuda.h>

#ifndef HMVIT_CHAIN_INC     // registers per thread of the transform warpgroups / of the TMA + MMA warpgroup after rebalancing
#define HMVIT_CHAIN_INC 104
#endif
#ifndef HMVIT_CHAIN_DEC
#define HMVIT_CHAIN_DEC 40
#endif
#ifndef HMVIT_CHAIN_DBG    // bottleneck-hunting builds only (results are wrong): 1 no stores, 2 no weight TMA, 4 no MMA, 8 no residual loads
#define HMVIT_CHAIN_DBG 0
#endif

namespace hmvit {

#ifdef HMVIT_TS   // timeline instrumentation build: CTA 0 records clock64() at phase boundaries of its first tiles
__device__ unsigned long long g_chain_ts[2][16][16];   // [role: 0 transform, 1 mma][tile][event]
#define CHAIN_TS(role, tile, ev) do { if (blockIdx.x == 0 && (tile) < 16) g_chain_ts[role][tile][ev] = clock64(); } while (0)
#else
#define CHAIN_TS(role, tile, ev) do { } while (0)
#endif

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
  int out_L;                   // agent slots per scene in out_cm (L; 1 for the head, whose output is [B][256][N])
  const float* hb1;            // fused head (mode 2): [2][256] biases and the [B][256][N] output
  const float* hb2;
  float* head_out;
};

struct ChainMaps {             // TMA tensor maps
  CUtensorMap o;               // attention output rows bf16 [B*L*N][256], box 64 x 128
  CUtensorMap wa[2];           // bf16 [256][256], box 64 x 256
  CUtensorMap w1[2];           // fp16 [256][256], box 64 x 256
  CUtensorMap w2[2];
  CUtensorMap hw1[2];          // fused head (mode 2): fp16 [256][256] head weights
  CUtensorMap hw2[2];
};

struct ChainCfg {
  static constexpr int BM = 128;
  static constexpr int CHUNK = 16384;                 // A chunk: 128 rows x 128 B
  static constexpr int WSTAGE = 32768;                // weight stage: 256 output channels x 128 B of K
  // NF must be 4: ring stage g belongs to transform group g (single producer per stage, generations in
  // program order -- an mbarrier parity wait cannot tell generation k from k - 2).
  static constexpr int NF = 4;                        // fp16 A-chunk ring stages (64 channels each)
  static constexpr int NS = 2;                        // weight ring stages (32 KB each)
  static constexpr int AO_BYTES = 4 * CHUNK;          // O tile, bf16
  static constexpr int PART_BYTES = 4 * 128 * 8;      // per-group partial LN statistics
  static constexpr int BIAS_BYTES = 5 * 2 * 256 * 4;   // ba | b1 | b2 | head b1 | head b2, both types: read by every transform thread
  static constexpr int SMEM_BYTES = AO_BYTES + NF * CHUNK + NS * WSTAGE + PART_BYTES + BIAS_BYTES + 256 + 1024;
  static constexpr int NT = 512;                      // transform threads
  static constexpr int THREADS = NT + 128;            // 16 transform warps + one warpgroup holding the TMA and MMA warps (2 idle)
};

// combine four (mean, M2) partials over 64 values each -> (mean, rstd) over 256
HMVIT_DEVINL float2 ln_combine(const float2* part, int row, float eps) {
  const float2 p0 = part[row], p1 = part[128 + row], p2 = part[256 + row], p3 = part[384 + row];
  const float mean = 0.25f * (p0.x + p1.x + p2.x + p3.x);
  const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
  const float m2 = p0.y + p1.y + p2.y + p3.y + 64.0f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
  return make_float2(mean, rsqrtf(fmaxf(m2 * (1.0f / kC), 0.f) + eps));
}

// kHead: the typed feed-forward HEAD of the fusion module (HeteroFusion.mlp_head, bevformer_point_pillar_hetero.py:36,
// 47-48): y = W_2 gelu(W_1 x + b_1) + b_2 on the ego rows -- the same pipeline without the output projection (P1),
// without LayerNorm and without the residual (P3 overwrites D1 instead of accumulating onto x').
// kMode 2: the normal chain on the ego tiles of the LAST stage followed, in the same tile, by the head: x'' is not
// stored but fed (fp16) into two more GEMM phases, P4: D2 = x'' Wh_1^T, P5: D1 = gelu(D2 + bh_1) Wh_2^T, and
// E3 writes D1 + bh_2 to the [B][256][N] output.  Saves the head's launch and its pipeline fill / drain (3.6 tiles per SM).
template <int kMode>
__global__ void __launch_bounds__(ChainCfg::THREADS, 1) chain_kernel(const __grid_constant__ ChainMaps maps, const ChainParams p) {
  using Cfg = ChainCfg;
  constexpr bool kHead = kMode == 1, kFuse = kMode == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sO = smem;
  uint8_t* sF = smem + Cfg::AO_BYTES;
  uint8_t* sW = sF + Cfg::NF * Cfg::CHUNK;
  float2* sPart = reinterpret_cast<float2*>(sW + Cfg::NS * Cfg::WSTAGE);  // [4 groups][128 rows]
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sPart) + Cfg::PART_BYTES);   // [3][2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sBias) + Cfg::BIAS_BYTES);
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
    mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d1_final, 1); mbar_init(d1_free, Cfg::NT);
    fence_mbar_init();
  }
  if (warp == 17) tmem_alloc<512>(tmem_slot);
  // biases in shared memory (a global load per use showed up as long-scoreboard stalls of the transform warps)
  for (int e = threadIdx.x; e < 2 * kC; e += Cfg::THREADS) {
    sBias[e] = __ldg(p.ba + e); sBias[2 * kC + e] = __ldg(p.b1 + e); sBias[4 * kC + e] = __ldg(p.b2 + e);
    if constexpr (kFuse) { sBias[6 * kC + e] = __ldg(p.hb1 + e); sBias[8 * kC + e] = __ldg(p.hb2 + e); }
  }
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

  if (warp < 16) {
    // ============================ transform / epilogue warps ============================
    // register rebalancing between warpgroups: the four transform warpgroups take 104 registers per thread,
    // the warpgroup of the TMA / MMA warps gives its share back (640 x 96 at launch : the pool is per CTA, 128 x 56 released >= 512 x 8 requested)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HMVIT_CHAIN_INC));
    const int gq = warp >> 2;                          // column / chunk group
    const int row = (warp & 3) * 32 + lane;            // token row == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int c0 = gq * 64;                            // first column owned in E1 / E2
    uint32_t ti = 0;
    auto next_valid = [&](int t) { int a_, k_; while (t < total_tiles && !tile_agent(t, a_, k_)) t += gridDim.x; return t; };
    // L2 prefetch of the NEXT tile's residual rows (256 channel segments of 512 B): issued while this
    // tile's FFN runs, so that the register loads at the start of the next tile hit L2 instead of HBM
    auto prefetch_resid_l2 = [&](int tn) {
      int a_, k_;
      if (threadIdx.x < kC && tn < total_tiles && tile_agent(tn, a_, k_)) {
        const float* r_ = p.resid_cm + (static_cast<size_t>(a_) * kC + threadIdx.x) * p.N + k_;
        const uint32_t bytes = static_cast<uint32_t>(min(Cfg::BM, p.N - k_)) * 4u;
        if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(r_) & 15u) == 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(r_), "r"(bytes) : "memory");
      }
    };
    int t = next_valid(blockIdx.x);
    while (t < total_tiles) {
      const int t_next = next_valid(t + gridDim.x);
      int a, tok0;
      tile_agent(t, a, tok0);
      const int type = p.mode[a] != 0 ? 1 : 0;
      const int tok = tok0 + row;
      const bool valid = tok < p.N;
      const size_t cm_off = static_cast<size_t>(a) * kC * p.N + (valid ? tok : 0);
      const float* res = p.resid_cm + cm_off + static_cast<size_t>(c0) * p.N;
      const int a_out = kHead ? (a / p.L) * p.out_L : a;        // head: slot 0 of scene b -> row b of the [B][256][N] output
      float* dst = p.out_cm + static_cast<size_t>(a_out) * kC * p.N + (valid ? tok : 0) + static_cast<size_t>(c0) * p.N;
      // ---- E1: x' = D1 + b_a + x on this group's 64 columns; the residual loads are issued before the
      //      projection MMAs have finished (and were L2-prefetched during the previous tile) ----
      float rv[64];
#pragma unroll
      for (int k = 0; k < 64; ++k) rv[k] = (valid && !(HMVIT_CHAIN_DBG & 8)) ? res[k * p.N] : 0.f;
      float rstd = 1.f, nmr = 0.f;
      if constexpr (!kHead) {
        if (threadIdx.x == 0) CHAIN_TS(0, ti, 0);
        mbar_wait(d1_full, ti & 1);
        tc_fence_after();
        if (threadIdx.x == 0) CHAIN_TS(0, ti, 1);
        float s0 = 0.f, sum = 0.f, sq = 0.f;
        {
          const float* bias = sBias + type * kC + c0;
  #pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t r[16];
            tmem_ld16(D1 + lane_base + c0 + q * 16, r);
            tmem_ld_wait();
  #pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float v = __uint_as_float(r[k]) + bias[q * 16 + k] + rv[q * 16 + k];
              if (q == 0 && k == 0) s0 = v;
              const float d = v - s0;
              sum += d; sq += d * d;
              r[k] = __float_as_uint(v);
              rv[q * 16 + k] = v;                    // x' stays in registers for the LayerNorm feed below
            }
            tmem_st16(D1 + lane_base + c0 + q * 16, r);
          }
        }
        tmem_st_wait();
        {
          const float md = sum * (1.0f / 64.0f);
          sPart[gq * 128 + row] = make_float2(s0 + md, fmaxf(sq - sum * md, 0.f));     // (mean_g, M2_g)
        }
        tc_fence_before();
        named_bar_sync(1, Cfg::NT);
        tc_fence_after();
        const float2 st = ln_combine(sPart, row, p.ln_eps);
        rstd = st.y; nmr = -st.x * st.y;
      }
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 2);
      // ---- P2 feed: LN'(x') -> fp16 K-chunk gq (this group's own 64 columns, from registers) ----
      const bool affine = p.ln_gamma != nullptr;
      const float* gam = affine ? p.ln_gamma + type * kC + c0 : nullptr;
      const float* bet = affine ? p.ln_beta + type * kC + c0 : nullptr;
      {
        // ring stage gq belongs to this group (single producer): 2 generations per tile (P2, P3)
        const uint32_t fs = gq;
        uint32_t r[32];
        if (affine) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            r[k] = pack_f16x2(fmaf(rv[2 * k], rstd, nmr) * __ldg(gam + 2 * k) + __ldg(bet + 2 * k),
                              fmaf(rv[2 * k + 1], rstd, nmr) * __ldg(gam + 2 * k + 1) + __ldg(bet + 2 * k + 1));
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) r[k] = pack_f16x2(fmaf(rv[2 * k], rstd, nmr), fmaf(rv[2 * k + 1], rstd, nmr));
        }
        mbar_wait(&f_empty[fs], 1u);                 // generation 2 ti: parity 0, wait on the previous one
        uint8_t* dstF = sF + fs * Cfg::CHUNK;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<uint4*>(dstF + sw128_offset(row, u)) = make_uint4(r[u * 4], r[u * 4 + 1], r[u * 4 + 2], r[u * 4 + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&f_full[fs]);
      }
      prefetch_resid_l2(t_next);
      // ---- P3 feed: gelu(D2 + b_1) -> fp16 K-chunk gq ----
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 3);
      mbar_wait(d2_full, kFuse ? 0u : (ti & 1));     // mode 2: two completions per tile (P2, P4)
      tc_fence_after();
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 4);
      const float* b1 = sBias + 2 * kC + type * kC + c0;
      {
        const uint32_t fs = gq;
        uint32_t h2[32];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t r[32];
          tmem_ld32(D2 + lane_base + c0 + hf * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k)
            h2[hf * 16 + k] = pack_f16x2(gelu_erf_fast(__uint_as_float(r[2 * k]) + b1[hf * 32 + 2 * k]),
                                         gelu_erf_fast(__uint_as_float(r[2 * k + 1]) + b1[hf * 32 + 2 * k + 1]));
        }
        mbar_wait(&f_empty[fs], 0u);                 // generation 2 ti + 1: the P2 chunk of this tile has been consumed
        uint8_t* dstF = sF + fs * Cfg::CHUNK;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          *reinterpret_cast<uint4*>(dstF + sw128_offset(row, u)) = make_uint4(h2[u * 4], h2[u * 4 + 1], h2[u * 4 + 2], h2[u * 4 + 3]);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&f_full[fs]);
      }
      // ---- E2: x'' = D1 + b_2 on this group's 64 columns (+ LayerNorm statistics of x'') ----
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 5);
      mbar_wait(d1_final, kFuse ? 0u : (ti & 1));    // mode 2: two completions per tile (P3, P5)
      tc_fence_after();
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 6);
      const float* b2 = sBias + 4 * kC + type * kC + c0;
      float t0 = 0.f, tsum = 0.f, tsq = 0.f;
      if constexpr (kFuse) {
        // ---- P4 feed: x'' = D1 + b_2 -> fp16 K-chunk gq (x'' is not stored: the head is its only consumer) ----
        {
          uint32_t r0[32], r1[32];
          tmem_ld32(D1 + lane_base + c0, r0);
          tmem_ld32(D1 + lane_base + c0 + 32, r1);
          tmem_ld_wait();
          uint32_t h2[32];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            h2[k] = pack_f16x2(__uint_as_float(r0[2 * k]) + b2[2 * k], __uint_as_float(r0[2 * k + 1]) + b2[2 * k + 1]);
            h2[16 + k] = pack_f16x2(__uint_as_float(r1[2 * k]) + b2[32 + 2 * k], __uint_as_float(r1[2 * k + 1]) + b2[32 + 2 * k + 1]);
          }
          mbar_wait(&f_empty[gq], 1u);               // generation 4 ti + 2
          uint8_t* dstF = sF + gq * Cfg::CHUNK;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(dstF + sw128_offset(row, u)) = make_uint4(h2[u * 4], h2[u * 4 + 1], h2[u * 4 + 2], h2[u * 4 + 3]);
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&f_full[gq]);
        }
        // ---- P5 feed: gelu(D2 + bh_1) -> fp16 K-chunk gq ----
        mbar_wait(d2_full, 1u);
        tc_fence_after();
        {
          const float* hb1 = sBias + 6 * kC + type * kC + c0;
          uint32_t h2[32];
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t r[32];
            tmem_ld32(D2 + lane_base + c0 + hf * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k)
              h2[hf * 16 + k] = pack_f16x2(gelu_erf_fast(__uint_as_float(r[2 * k]) + hb1[hf * 32 + 2 * k]),
                                           gelu_erf_fast(__uint_as_float(r[2 * k + 1]) + hb1[hf * 32 + 2 * k + 1]));
          }
          mbar_wait(&f_empty[gq], 0u);               // generation 4 ti + 3
          uint8_t* dstF = sF + gq * Cfg::CHUNK;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            *reinterpret_cast<uint4*>(dstF + sw128_offset(row, u)) = make_uint4(h2[u * 4], h2[u * 4 + 1], h2[u * 4 + 2], h2[u * 4 + 3]);
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(&f_full[gq]);
        }
        // ---- E3: head output = D1 + bh_2 on this group's 64 columns ----
        mbar_wait(d1_final, 1u);
        tc_fence_after();
        {
          const float* hb2 = sBias + 8 * kC + type * kC + c0;
          float* hdst = p.head_out + static_cast<size_t>(a / p.L) * kC * p.N + static_cast<size_t>(c0) * p.N + (valid ? tok : 0);
          uint32_t r0[32], r1[32];
          tmem_ld32(D1 + lane_base + c0, r0);
          tmem_ld32(D1 + lane_base + c0 + 32, r1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(d1_free);                      // D1 is in registers: the next tile's projection may start
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (valid) hdst[k * p.N] = __uint_as_float(k < 32 ? r0[k] : r1[k - 32]) + hb2[k];
        }
      } else
      {
        uint32_t r0[32], r1[32];
        tmem_ld32(D1 + lane_base + c0, r0);
        tmem_ld32(D1 + lane_base + c0 + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(d1_free);                      // D1 is in registers: the next tile's projection may start
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const float v = __uint_as_float(k < 32 ? r0[k] : r1[k - 32]) + b2[k];
          if (k == 0) t0 = v;
          const float d = v - t0;
          tsum += d; tsq += d * d;
          if (valid && (!(HMVIT_CHAIN_DBG & 1) || v == 12345.678f)) dst[k * p.N] = v;
        }
      }
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 7);
      if (!kFuse && p.stats_out != nullptr) {
        const float tmd = tsum * (1.0f / 64.0f);
        sPart[gq * 128 + row] = make_float2(t0 + tmd, fmaxf(tsq - tsum * tmd, 0.f));
        named_bar_sync(1, Cfg::NT);
        if (gq == 0 && valid) p.stats_out[static_cast<size_t>(a) * p.N + tok] = ln_combine(sPart, row, p.ln_eps);
        named_bar_sync(1, Cfg::NT);     // sPart is rewritten by the next tile's E1
      }
      if (threadIdx.x == 0) CHAIN_TS(0, ti, 8);
      ++ti;
      t = t_next;
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HMVIT_CHAIN_DEC));     // one instruction for the whole warpgroup (warps 16-19)
  if (warp == 16) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      if constexpr (!kHead) tma_prefetch_desc(&maps.o);
      uint32_t ti = 0, itw = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int a, tok0;
        if (!tile_agent(t, a, tok0)) continue;
        const int type = p.mode[a] != 0 ? 1 : 0;
        if constexpr (!kHead) {
          mbar_wait(o_empty, (ti & 1) ^ 1u);
          mbar_arrive_expect_tx(o_full, Cfg::AO_BYTES);
          const int row0 = a * p.N + tok0;
          for (int kc = 0; kc < 4; ++kc) tma_load_2d(sO + kc * Cfg::CHUNK, &maps.o, o_full, kc * 64, row0);
        }
        auto wstage = [&](const CUtensorMap* m, int c0) {
          const uint32_t s = itw % Cfg::NS, ph = (itw / Cfg::NS) & 1u;
          mbar_wait(&w_empty[s], ph ^ 1u);
          if ((HMVIT_CHAIN_DBG & 2) && itw >= Cfg::NS) { mbar_arrive(&w_full[s]); ++itw; return; }
          mbar_arrive_expect_tx(&w_full[s], Cfg::WSTAGE);
          tma_load_2d(sW + s * Cfg::WSTAGE, m, &w_full[s], c0, 0);
          ++itw;
        };
        if constexpr (!kHead) for (int kc = 0; kc < 4; ++kc) wstage(&maps.wa[type], kc * 64);
        for (int j = 0; j < 4; ++j) wstage(&maps.w1[type], j * 64);   // K-chunk j == transform group j, same order as the MMA issuer
        for (int j = 0; j < 4; ++j) wstage(&maps.w2[type], j * 64);
        if constexpr (kFuse) {
          for (int j = 0; j < 4; ++j) wstage(&maps.hw1[type], j * 64);
          for (int j = 0; j < 4; ++j) wstage(&maps.hw2[type], j * 64);
        }
        ++ti;
      }
    }
  } else if (warp == 17) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      constexpr uint32_t idesc_bf16 = umma_idesc(1u, 128, 256);
      constexpr uint32_t idesc_f16 = umma_idesc(0u, 128, 256);     // fp16 operands, fp32 accumulate
      const uint32_t o_base = smem_u32(sO), f_base = smem_u32(sF), w_base = smem_u32(sW);
      uint32_t ti = 0, itw = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int a, tok0;
        if (!tile_agent(t, a, tok0)) continue;
        CHAIN_TS(1, ti, 0);
        mbar_wait(d1_free, (ti & 1) ^ 1u);
        CHAIN_TS(1, ti, 1);
        if constexpr (!kHead) {
          mbar_wait(o_full, ti & 1);
          tc_fence_after();
          CHAIN_TS(1, ti, 2);
          // P1: D1 = O W_a^T   (4 weight stages of 64 K-columns, one M128 N256 K16 MMA per k-step)
          for (int kc = 0; kc < 4; ++kc, ++itw) {
            const uint32_t s = itw % Cfg::NS, ph = (itw / Cfg::NS) & 1u;
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
            if (!(HMVIT_CHAIN_DBG & 4))
  #pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss<2>(D1, umma_desc_sw128(o_base + kc * Cfg::CHUNK + ks * 32),
                         umma_desc_sw128(w_base + s * Cfg::WSTAGE + ks * 32), idesc_bf16, (kc | ks) != 0 ? 1u : 0u);
            umma_commit(&w_empty[s]);
          }
          umma_commit(o_empty);
          umma_commit(d1_full);
        }
        CHAIN_TS(1, ti, 3);
        // P2: D2 = LN'(x') W_1^T      P3: D1 += gelu(.) W_2^T     (4 stages of 64 K-columns each)
        for (int phase = 0; phase < (kFuse ? 4 : 2); ++phase) {   // mode 2: + P4 (D2 = x'' Wh_1^T) and P5 (D1 = gelu(.) Wh_2^T)
          const uint32_t dacc = (phase & 1) == 0 ? D2 : D1;
          for (int j = 0; j < 4; ++j, ++itw) {
            const uint32_t fs = j, fph = phase & 1;              // generation (2 | 4) ti + phase of ring stage j
            const uint32_t s = itw % Cfg::NS, ph = (itw / Cfg::NS) & 1u;
            mbar_wait(&f_full[fs], fph);
            mbar_wait(&w_full[s], ph);
            tc_fence_after();
            if (!(HMVIT_CHAIN_DBG & 4))
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_ss<2>(dacc, umma_desc_sw128(f_base + fs * Cfg::CHUNK + ks * 32),
                         umma_desc_sw128(w_base + s * Cfg::WSTAGE + ks * 32), idesc_f16,
                         ((phase == 1 && !kHead) || (j | ks) != 0) ? 1u : 0u);   // P3 accumulates onto x'; P4 / P5 overwrite
            umma_commit(&w_empty[s]);
            umma_commit(&f_empty[fs]);
          }
          umma_commit((phase & 1) == 0 ? d2_full : d1_final);
          if (phase < 2) CHAIN_TS(1, ti, 4 + phase);
        }
        ++ti;
      }
    }
  }

  }

  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace hmvit
