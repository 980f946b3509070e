// Dense multi-agent group attention on the 5th-gen tensor cores (tcgen05 + TMEM): the second launch of the split
// attention (attn_split.cuh), same contract as dense_attn_kernel -- replaces the per-ego HeteroAttention.forward
// core (hetero_fusion.py:187-277) on the visible, already warped + blended key / value tiles written by
// warp_compact_kernel.
//
// One CTA per (scene b, ego i, token group g, head group of 4 heads) = two head PAIRS; 10 warps:
//   warps 0-3   softmax of pair 0   thread == TMEM lane == (head of the pair, query row): S from TMEM, + relative
//   warps 4-7   softmax of pair 1   position bias (looked up through the tile's key-slot list), running max, exp2,
//                                   bf16 P row written back to TMEM (over S), D rescaled in TMEM when the max moved
//   warp  8     MMA issuer          converged warp, one elected lane:  S = Qbd K^T  (M128 N64 K64, block-diagonal Q
//                                   stacks the pair's two heads on the M axis),  D += P V  (P from TMEM, V MN-major)
//   warp  9     producer            per 64-key tile two 16 KB bulk copies (K, V: UMMA SWIZZLE_128B images written by
//                                   the compaction pass) + the tile's bias offsets, double buffered
// Why tcgen05 here: the mma.sync form re-reads every K / V fragment from shared memory once per 16-row block
// (ldmatrix) and was bound by the shared-memory port (~70 % busy, profiles/); UMMA reads each operand tile once.
// The two pairs ping-pong: while one pair's softmax runs, the tensor core serves the other pair.
#pragma once
#include "attn_split.cuh"
#include "attn_tc.cuh"

#ifndef HMVIT_DTC_DBG   // bottleneck-hunting builds only (results are wrong): 1 no softmax math, 2 no tile copies, 4 no MMAs
#define HMVIT_DTC_DBG 0
#endif

namespace hmvit {

#ifdef HMVIT_TS   // timeline instrumentation: CTAs (x < 8, y == 0) record clock64() per role (tools/dense_tc_timeline.py)
__device__ unsigned long long g_dtc_ts[8][4][32];   // [cta][role: 0 softmax pair 0, 1 softmax pair 1, 2 mma, 3 producer][event]
#define DTC_TS(role, idx) do { if (blockIdx.y == 0 && blockIdx.x < 8 && (idx) < 32) g_dtc_ts[blockIdx.x][role][idx] = clock64(); } while (0)
#else
#define DTC_TS(role, idx) do { } while (0)
#endif

struct DenseTcCfg {
  static constexpr int THREADS = 320;
  static constexpr int OFF_Q = 0;                       // [2 pairs][128 rows][128 B]  block-diagonal Q
  static constexpr int OFF_KV = 32768;                  // [2 buffers][K: 2 pairs x 8 KB | V: 2 pairs x 8 KB]
  static constexpr int KV_BUF = 32768;
  static constexpr int OFF_BIAS = OFF_KV + 2 * KV_BUF;  // [4 heads][kBiasStride] fp32, log2 domain
  static constexpr int OFF_KOFF = OFF_BIAS + kHG * kBiasStride * 4;   // [2 buffers][64] bias offset of every key (uint8)
  static constexpr int OFF_BAR = OFF_KOFF + 128;
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr uint32_t TM_COLS = 256;              // pair p: S / P at 128 p, D at 128 p + 64
};

__global__ void __launch_bounds__(DenseTcCfg::THREADS, 2) dense_attn_tc_kernel(const SplitParams sp) {
  using Cfg = DenseTcCfg;
  const AttnParams& p = sp.a;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = p.record_len[b];
  if (i >= nrec || (p.ego_only && i != 0)) return;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int grp = blockIdx.x >> 1, hgc = blockIdx.x & 1;
  const int gy = grp / GX, gx = grp - gy * GX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sKV = smem + Cfg::OFF_KV;
  float* sBias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint8_t* sKoff = smem + Cfg::OFF_KOFF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars + 0;      // [2]
  uint64_t* kv_empty = bars + 2;     // [2]
  uint64_t* s_full = bars + 4;       // [2 pairs]
  uint64_t* p_full = bars + 6;       // [2 pairs]
  uint64_t* d_full = bars + 8;       // [2 pairs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const size_t ag = static_cast<size_t>(a) * G + grp;
  const int nv = sp.nvis[ag];
  const int ntiles = (nv + kS - 1) >> 6;
  const int self = self_is_identity(p, b, i) ? 1 : 0;
  const int nvt = self + ntiles;                        // key tiles: [self tile] [compacted tiles 0..]
  const int cu0 = hgc * 16;                             // first 16-byte unit of this head group in a 512-byte row

  if (tid == 0) DTC_TS(0, 0);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); mbar_init(&d_full[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<Cfg::TM_COLS>(tmem_slot);

  // ------------------------------ one-off staging (softmax warps) ------------------------------
  if (warp < 8) {
    const int te = p.mode[a] != 0 ? 1 : 0;
    const int hl = lane >> 4, u16 = lane & 15;
    const int pr = u16 >> 3, hq = (u16 >> 2) & 1, cq = u16 & 3, un = u16 & 7;
    const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    const uint4* ksrc = reinterpret_cast<const uint4*>(p.k) + te * plane + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    const uint4* vsrc = reinterpret_cast<const uint4*>(p.v) + te * plane + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    uint4 qv[4], kv[4], vv[4];
#pragma unroll
    for (int tt = 0; tt < 4; ++tt) {
      const int s = warp * 8 + tt * 2 + hl;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      const size_t off = static_cast<size_t>(r * p.W + c) * 32;
      qv[tt] = __ldg(qsrc + off);
      if (self) { kv[tt] = __ldg(ksrc + off); vv[tt] = __ldg(vsrc + off); }
    }
    for (int e = tid; e < 225 * kHG; e += 256)
      sBias[(e & 3) * kBiasStride + (e >> 2)] = __ldg(p.bias_table + (e >> 2) * kHeads + hgc * kHG + (e & 3)) * 1.4426950408889634f;
    uint32_t bk2[4] = {0, 0, 0, 0}, bv2[4] = {0, 0, 0, 0};
    if (self) {
      // own keys / values: one tap of weight 1 on top of the folded bias -- the arithmetic of the general gather
      const float4* pk = reinterpret_cast<const float4*>(p.bk + (te * 2 + te) * kC + (cu0 + u16) * 8);
      const float4* pv = reinterpret_cast<const float4*>(p.bv + (te * 2 + te) * kC + (cu0 + u16) * 8);
      const float4 k0 = __ldg(pk), k1 = __ldg(pk + 1), v0 = __ldg(pv), v1 = __ldg(pv + 1);
      bk2[0] = pack_bf16x2(k0.x, k0.y); bk2[1] = pack_bf16x2(k0.z, k0.w); bk2[2] = pack_bf16x2(k1.x, k1.y); bk2[3] = pack_bf16x2(k1.z, k1.w);
      bv2[0] = pack_bf16x2(v0.x, v0.y); bv2[1] = pack_bf16x2(v0.z, v0.w); bv2[2] = pack_bf16x2(v1.x, v1.y); bv2[3] = pack_bf16x2(v1.z, v1.w);
      if (tid < 64) sKoff[tid] = static_cast<uint8_t>((tid >> 3) * 15 + (tid & 7));
    }
    constexpr uint32_t kOne2 = 0x3F803F80u;
#pragma unroll
    for (int tt = 0; tt < 4; ++tt) {
      const int s = warp * 8 + tt * 2 + hl;
      // block-diagonal Q: pair pr, rows [0,64) hold head 2pr in K-columns [0,32), rows [64,128) head 2pr+1 in [32,64)
      *reinterpret_cast<uint4*>(sQ + pr * 16384 + sw128_offset(hq * 64 + s, hq * 4 + cq)) = qv[tt];
      *reinterpret_cast<uint4*>(sQ + pr * 16384 + sw128_offset((1 - hq) * 64 + s, hq * 4 + cq)) = make_uint4(0, 0, 0, 0);
      if (self) {
        uint4 ko, vo;
        ko.x = hfma2_bf16(kOne2, kv[tt].x, bk2[0]); ko.y = hfma2_bf16(kOne2, kv[tt].y, bk2[1]);
        ko.z = hfma2_bf16(kOne2, kv[tt].z, bk2[2]); ko.w = hfma2_bf16(kOne2, kv[tt].w, bk2[3]);
        vo.x = hfma2_bf16(kOne2, vv[tt].x, bv2[0]); vo.y = hfma2_bf16(kOne2, vv[tt].y, bv2[1]);
        vo.z = hfma2_bf16(kOne2, vv[tt].z, bv2[2]); vo.w = hfma2_bf16(kOne2, vv[tt].w, bv2[3]);
        *reinterpret_cast<uint4*>(sKV + pr * 8192 + sw128_offset(s, un)) = ko;
        *reinterpret_cast<uint4*>(sKV + 16384 + pr * 8192 + sw128_offset(s, un)) = vo;
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  if (tid == 0) DTC_TS(0, 1);

  if (warp < 8) {
    // =========================================== SOFTMAX ===========================================
    const int pr = warp >> 2;                                   // head pair of this warpgroup
    const int L7 = tid & 127, hh = L7 >> 6, row = L7 & 63;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tm + lane_base + pr * 128, tD = tS + 64 + hh * 32;
    const float* bt = sBias + (pr * 2 + hh) * kBiasStride + ((row >> 3) + 7) * 15 + (row & 7) + 7;   // bt[-koff(key slot)]
    float m_run = -INFINITY, l_run = 0.f;
    for (int vt = 0; vt < nvt; ++vt) {
      const int nval = vt < self ? kS : min(kS, nv - (vt - self) * kS);
      const uint8_t* ko = sKoff + (vt & 1) * 64;
      if ((tid & 127) == 0) DTC_TS(pr, 2 + vt * 3);
      mbar_wait(&s_full[pr], vt & 1);
      tc_fence_after();
      if ((tid & 127) == 0) DTC_TS(pr, 3 + vt * 3);
      // the 64 logits of this row are processed as two halves of 32 keys; the first half (bias added) is parked in
      // its own TMEM columns while the second one is reduced, so that at most 32 logits are live in registers
      float mx = -INFINITY;
      auto add_bias = [&](uint32_t (&sv)[32], int j0) {
        if (HMVIT_DTC_DBG & 1) { mx = 0.f; return; }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 kw = *reinterpret_cast<const uint4*>(ko + j0 + q * 16);      // 16 key offsets (broadcast)
          const uint32_t w4[4] = {kw.x, kw.y, kw.z, kw.w};
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int kf = (w4[e >> 2] >> ((e & 3) * 8)) & 0xff;
            const float v = __uint_as_float(sv[q * 16 + e]) + bt[-kf];
            sv[q * 16 + e] = __float_as_uint(v);
          }
        }
        if (j0 + 32 > nval) {                                                      // tail of the last tile
#pragma unroll
          for (int e = 0; e < 32; ++e) if (j0 + e >= nval) sv[e] = 0xff800000u;      // -inf
        }
#pragma unroll
        for (int e = 0; e < 32; e += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));   // FMNMX3
      };
      uint32_t pk[32];
      float alpha, lsum = 0.f;
      {
        uint32_t sa[32];
        tmem_ld32(tS, sa);
        tmem_ld_wait();
        add_bias(sa, 0);
        tmem_st32(tS, sa);
      }
      {
        uint32_t sb[32];
        tmem_ld32(tS + 32, sb);
        tmem_ld_wait();
        add_bias(sb, 32);
        const float m_new = fmaxf(m_run, mx);
        const float mu = (m_new == -INFINITY) ? 0.f : m_new;
        alpha = ex2(m_run - mu);                                 // 0 for the first tile (m_run = -inf)
        m_run = m_new;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (HMVIT_DTC_DBG & 1) { pk[16 + k] = sb[k]; continue; }
          pk[16 + k] = pack_bf16x2(ex2(__uint_as_float(sb[2 * k]) - mu), ex2(__uint_as_float(sb[2 * k + 1]) - mu));
          lsum += bf16_lo(pk[16 + k]) + bf16_hi(pk[16 + k]);     // denominator from exactly the bf16 probabilities that multiply V
        }
        uint32_t sa[32];
        tmem_st_wait();
        tmem_ld32(tS, sa);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (HMVIT_DTC_DBG & 1) { pk[k] = sa[k]; continue; }
          pk[k] = pack_bf16x2(ex2(__uint_as_float(sa[2 * k]) - mu), ex2(__uint_as_float(sa[2 * k + 1]) - mu));
          lsum += bf16_lo(pk[k]) + bf16_hi(pk[k]);
        }
      }
      l_run = l_run * alpha + lsum;
      if (vt > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) {
        // the running max moved: rescale this head's accumulator in TMEM (P V of the previous tile has retired:
        // the commit behind s_full covers every earlier MMA)
        uint32_t d[32];
        tmem_ld32(tD, d);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) d[k] = __float_as_uint(__uint_as_float(d[k]) * alpha);
        tmem_st32(tD, d);
      }
      tmem_st32(tS, pk);                                         // P (bf16 pairs, A operand of P V) over the consumed logits
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[pr]);
      if ((tid & 127) == 0) DTC_TS(pr, 4 + vt * 3);
    }
    if ((tid & 127) == 0) DTC_TS(pr, 26);
    // ------------------------------ normalise and store ------------------------------
    int r, c; group_token(p.kind, gy, gx, row, p.H, p.W, r, c);
    const size_t tok = static_cast<size_t>(a) * N + r * p.W + c;
    const int head = hgc * kHG + pr * 2 + hh;
    uint32_t d[32];
    if (nvt > 0) {
      mbar_wait(&d_full[pr], 0);
      tc_fence_after();
      tmem_ld32(tD, d);
      tmem_ld_wait();
    }
    const float il = (nvt > 0 && l_run > 0.f) ? 1.0f / l_run : 0.f;
    uint4* dst = reinterpret_cast<uint4*>(p.out + tok * kC + head * kDh);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint4 o = make_uint4(0, 0, 0, 0);
      if (il != 0.f) {
        o.x = pack_bf16x2(__uint_as_float(d[u * 8 + 0]) * il, __uint_as_float(d[u * 8 + 1]) * il);
        o.y = pack_bf16x2(__uint_as_float(d[u * 8 + 2]) * il, __uint_as_float(d[u * 8 + 3]) * il);
        o.z = pack_bf16x2(__uint_as_float(d[u * 8 + 4]) * il, __uint_as_float(d[u * 8 + 5]) * il);
        o.w = pack_bf16x2(__uint_as_float(d[u * 8 + 6]) * il, __uint_as_float(d[u * 8 + 7]) * il);
      }
      dst[u] = o;
    }
    // training: softmax statistics (log2 domain: running max + log2 of the denominator)
    if (p.lse != nullptr) p.lse[tok * kHeads + head] = l_run > 0.f ? m_run + log2f(l_run) : INFINITY;
    if ((tid & 127) == 0) DTC_TS(pr, 27);
  } else if (warp == 8) {
    // =========================================== MMA ===========================================
    constexpr uint32_t idesc_qk = umma_idesc(1u, 128, 64);
    constexpr uint32_t idesc_pv = umma_idesc(1u, 128, 64) | (1u << 16);        // B (V tile) MN-major
    const uint32_t q_u = smem_u32(sQ), kv_u = smem_u32(sKV);
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    auto issue_qk = [&](int pr, int buf) {                       // S_pr = Qbd_pr K_pr^T
      if (HMVIT_DTC_DBG & 4) return;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_ss<2>(tmu + pr * 128, umma_desc_sw128(q_u + pr * 16384 + ks * 32),
                   umma_desc_sw128(kv_u + buf * Cfg::KV_BUF + pr * 8192 + ks * 32), idesc_qk, ks != 0 ? 1u : 0u);
    };
    auto issue_pv = [&](int pr, int buf, bool first) {           // D_pr (+)= P_pr V_pr
      if (HMVIT_DTC_DBG & 4) return;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_ts_bf16(tmu + pr * 128 + 64, tmu + pr * 128 + ks * 8,
                     umma_desc_sw128_mn(kv_u + buf * Cfg::KV_BUF + 16384 + pr * 8192 + ks * 2048), idesc_pv,
                     (!first || ks != 0) ? 1u : 0u);
    };
    if (nvt > 0) {
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(0, 0); umma_commit(&s_full[0]);
        issue_qk(1, 0); umma_commit(&s_full[1]);
      }
      __syncwarp();
      for (int vt = 0; vt < nvt; ++vt) {
        const int buf = vt & 1;
        const bool has_next = vt + 1 < nvt;
        if (has_next) { mbar_wait(&kv_full[buf ^ 1], ((vt + 1) >> 1) & 1); tc_fence_after(); }
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          mbar_wait(&p_full[pr], vt & 1);
          tc_fence_after();
          if (elect_one()) {
            issue_pv(pr, buf, vt == 0);
            if (has_next) { issue_qk(pr, buf ^ 1); umma_commit(&s_full[pr]); }
            else umma_commit(&d_full[pr]);
            if (pr == 1) umma_commit(&kv_empty[buf]);            // both pairs' P V of this tile issued
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =========================================== PRODUCER ===========================================
    if (self && lane == 0) mbar_arrive(&kv_full[0]);             // the self tile was staged above
    const uint8_t* kblob = sp.kc + ag * (static_cast<size_t>(p.L) * 2 * kBlobBytes) + hgc * kBlobBytes;
    const uint8_t* vblob = sp.vc + ag * (static_cast<size_t>(p.L) * 2 * kBlobBytes) + hgc * kBlobBytes;
    const uint8_t* slots = sp.slots + ag * (p.L * kS);
    const uint32_t kv_u = smem_u32(sKV);
    for (int vt = self; vt < nvt; ++vt) {
      const int tile = vt - self, buf = vt & 1;
      if (vt >= 2) mbar_wait(&kv_empty[buf], ((vt >> 1) - 1) & 1);     // P V of tile vt - 2 has retired
      {
        const uint32_t s2 = __ldg(reinterpret_cast<const uint16_t*>(slots + tile * kS) + lane);   // two key slots
        const uint32_t k2 = ((s2 >> 3) & 0x0707u) * 15u + (s2 & 0x0707u);                          // per byte, <= 112: no carry
        reinterpret_cast<uint16_t*>(sKoff + buf * 64)[lane] = static_cast<uint16_t>(k2);
      }
      __syncwarp();
      if ((HMVIT_DTC_DBG & 2) && lane == 0) mbar_arrive(&kv_full[buf]);
      if (!(HMVIT_DTC_DBG & 2) && lane == 0) {
        mbar_arrive_expect_tx(&kv_full[buf], 2 * kBlobBytes);
        bulk_load(kv_u + buf * Cfg::KV_BUF, kblob + static_cast<size_t>(tile) * 2 * kBlobBytes, kBlobBytes, &kv_full[buf]);
        bulk_load(kv_u + buf * Cfg::KV_BUF + 16384, vblob + static_cast<size_t>(tile) * 2 * kBlobBytes, kBlobBytes, &kv_full[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid == 0) DTC_TS(0, 28);
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TM_COLS>(tm);
    if (lane == 0) DTC_TS(2, 29);
  }
}

}  // namespace hmvit
