// Fused multi-agent group attention on the 5th-gen tensor cores (tcgen05 + TMEM), one CTA per
// (scene b, ego agent i, token group g, head group of 4 heads).
//
// Same contract as group_attn_kernel (attn.cuh) -- replaces warp_features + the per-ego
// HeteroAttention.forward core (hetero_fusion.py:338-361, 187-277) -- but warp-specialised:
//
//   warps 4-7  GATHER   per source agent j: 4-tap bilinear gather of the projected K' / V' rows (16-byte
//                       loads kept in flight in registers, packed-bf16 blend on top of the folded bias)
//                       into UMMA operand tiles; invisible keys get all-zero K / V rows and a 0 in the
//                       visibility row that feeds the softmax denominator
//   warp  8    MMA      one thread issues tcgen05.mma (M128, two heads stacked on the M axis through a
//                       block-diagonal Q tile):  S = Qbd K^T (N64),  D += P V (N64, V tile MN-major),
//                       Lsum += P vis (N16)
//   warps 0-3  SOFTMAX  thread == TMEM lane == (head of the pair, query row): S from TMEM, + relative
//                       position bias (128-bit loads from 4 pre-shifted copies of the table), running
//                       max, exp2, bf16 P row -> smem; rescales D / Lsum in TMEM when the max moved
//
// The key mask costs no instruction in the softmax: masked keys have K = 0 (their logit is just the
// bias, finite), V = 0 and vis = 0, so they add nothing to the numerator nor to the denominator
// (the denominator is the tensor-core product P x vis of exactly the bf16 probabilities that multiply V).
// Sources with no visible key in the group are skipped (softmax weight exactly 0).
#pragma once
#include "attn.cuh"

namespace hmvit {

struct AttnTcCfg {
  static constexpr int kMaxSrc = 8;                 // sources per tap pass
  static constexpr int THREADS = 288;
  static constexpr int OFF_Q = 0;                   // [2 pairs][128 rows][128 B]  block-diagonal Q
  static constexpr int OFF_K = 32768;               // [2 pairs][64 keys][128 B]   K-major
  static constexpr int OFF_V = 49152;               // [2 stages][2 pairs][64 keys][128 B]   MN-major (row = key)
  static constexpr int OFF_VIS = 81920;             // [2 stages][16][128 B]       row 0 = key visibility (bf16 1 / 0)
  static constexpr int OFF_BIAS = 86016;            // [4 shifts][4 heads][232] fp32, reversed table, log2 domain
  static constexpr int BIAS_STRIDE = 232;
  static constexpr int OFF_TAP = OFF_BIAS + 4 * 4 * BIAS_STRIDE * 4;         // [kMaxSrc][64] TapRec
  static constexpr int OFF_MAP = OFF_TAP + kMaxSrc * kS * 12;                // [kMaxSrc] source-pixel maps (+ valid flag)
  static constexpr int OFF_KVB = OFF_MAP + kMaxSrc * 80;                     // [kMaxSrc][2 (K, V)][128 ch] bf16 folded biases
  static constexpr int OFF_MISC = OFF_KVB + kMaxSrc * 512;                   // flags, barriers, TMEM slot
  static constexpr int SMEM_BYTES = OFF_MISC + 256 + 1024;                   // + alignment slack
  // TMEM columns: S logits, D / Lsum accumulators of the two head pairs, P probabilities (bf16 pairs, A operand)
  static constexpr uint32_t TM_S = 0, TM_D0 = 64, TM_P = 128, TM_L0 = 160, TM_L1 = 176, TM_D1 = 192, TM_COLS = 256;
};

struct SrcMap { WarpMap wm; int valid; int pad; };   // 80 bytes

#ifndef HMVIT_TC_DBG   // bottleneck-hunting builds only (results are wrong): 1 no tap loads, 2 no softmax math, 4 no L2 prefetch
#define HMVIT_TC_DBG 0
#endif

#ifdef HMVIT_TS   // timeline instrumentation: first 8 CTAs of agent 0 record clock64() per role
__device__ unsigned long long g_tc_ts[8][3][64];
#define TC_TS(role, idx) do { if (blockIdx.y == 0 && blockIdx.x < 8 && (idx) < 64) g_tc_ts[blockIdx.x][role][idx] = clock64(); } while (0)
#else
#define TC_TS(role, idx) do { } while (0)
#endif

HMVIT_DEVINL void tmem_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}
HMVIT_DEVINL void tmem_st1(uint32_t taddr, uint32_t r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r) : "memory");
}
// MN-major operand (rows of 128 B = 64 consecutive MN elements for one k; 8-row groups 1024 B apart)
HMVIT_DEVINL uint64_t umma_desc_sw128_mn(uint32_t smem_addr) { return umma_desc_sw128(smem_addr); }
// folded key / value biases of the pass's sources for this head group, as bf16 rows in shared memory
HMVIT_DEVINL void stage_kv_bias(const AttnParams& p, uint8_t* sKvb, int b, int j0, int nsrc, int te, int hgc, int t, int nt) {
  for (int e = t; e < nsrc * 2 * 64; e += nt) {              // one bf16 pair per element
    const int js = e >> 7, kv = (e >> 6) & 1, c2 = e & 63;
    const int tj = p.mode[b * p.L + j0 + js] != 0 ? 1 : 0;
    const float2 v = __ldg(reinterpret_cast<const float2*>((kv == 0 ? p.bk : p.bv) + (te * 2 + tj) * kC + hgc * 128) + c2);
    reinterpret_cast<uint32_t*>(sKvb)[e] = pack_bf16x2(v.x, v.y);
  }
}

__global__ void __launch_bounds__(AttnTcCfg::THREADS, 2) group_attn_tc_kernel(const AttnParams p) {
  using Cfg = AttnTcCfg;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = p.record_len[b];
  if (i >= nrec || (p.ego_only && i != 0)) return;
  const int N = p.H * p.W;
  const int GX = p.W / kWin;
  const int grp = blockIdx.x >> 1, hgc = blockIdx.x & 1;
  const int gy = grp / GX, gx = grp - gy * GX;
  const int te = p.mode[a] != 0 ? 1 : 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sK = smem + Cfg::OFF_K;
  uint8_t* sV = smem + Cfg::OFF_V;
  uint8_t* sVis = smem + Cfg::OFF_VIS;
  SrcMap* sMap = reinterpret_cast<SrcMap*>(smem + Cfg::OFF_MAP);
  uint8_t* sKvb = smem + Cfg::OFF_KVB;
  float* sBias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  TapRec* sTapAll = reinterpret_cast<TapRec*>(smem + Cfg::OFF_TAP);
  int* sAnyVis = reinterpret_cast<int*>(smem + Cfg::OFF_MISC);                 // [kMaxSrc]
  int* sAct = sAnyVis + Cfg::kMaxSrc;                                          // [kMaxSrc] active sources of the pass, [kMaxSrc] = count
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_MISC + 128);
  uint64_t* k_full = bars + 0;
  uint64_t* k_empty = bars + 1;
  uint64_t* v_full = bars + 2;      // [2]
  uint64_t* v_empty = bars + 4;     // [2]
  uint64_t* s_full = bars + 6;
  uint64_t* s_free = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* p_empty = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int cu0 = hgc * 16;                         // first 16-byte unit of this head group in a 512-byte row
  if (tid == 0) TC_TS(2, 0);

  // ------------------------------ one-off staging ------------------------------
  if (tid == 0) {
    mbar_init(k_full, 128); mbar_init(k_empty, 1);
    for (int st = 0; st < 2; ++st) { mbar_init(&v_full[st], 128); mbar_init(&v_empty[st], 1); }
    mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128); mbar_init(p_empty, 1);
    fence_mbar_init();
  }
  if (warp == 8) {
    tmem_alloc<Cfg::TM_COLS>(tmem_slot);
    for (int e = lane; e < 256; e += 32) *reinterpret_cast<uint4*>(sVis + e * 16) = make_uint4(0, 0, 0, 0);
    // source-pixel maps of the first tap pass (fp64, once per source instead of once per token)
    if (lane < min(Cfg::kMaxSrc, nrec)) {
      sMap[lane].valid = p.cav_mask[b * p.L + lane] != 0 ? 1 : 0;
      sMap[lane].wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + lane) * p.L + i) * 16, p.H, p.W, p.cell);
    }
    stage_kv_bias(p, sKvb, b, 0, min(Cfg::kMaxSrc, nrec), te, hgc, lane, 32);
  } else if (warp >= 4) {
    // block-diagonal Q tiles: pair pr, rows [0,64) hold head 2pr in K-columns [0,32), rows [64,128) hold
    // head 2pr+1 in K-columns [32,64); the other halves are zero
    const int gw = warp - 4, hl = lane >> 4, u16 = lane & 15;
    const int pr = u16 >> 3, hq = (u16 >> 2) & 1, cq = u16 & 3;
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
#pragma unroll 4
    for (int tt = 0; tt < 8; ++tt) {
      const int s = gw * 16 + tt * 2 + hl;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      const uint4 v = __ldg(qsrc + static_cast<size_t>(r * p.W + c) * 32);
      *reinterpret_cast<uint4*>(sQ + pr * 16384 + sw128_offset(hq * 64 + s, hq * 4 + cq)) = v;
      *reinterpret_cast<uint4*>(sQ + pr * 16384 + sw128_offset((1 - hq) * 64 + s, hq * 4 + cq)) = make_uint4(0, 0, 0, 0);
    }
  } else {
    // relative position bias: reversed table (so that keys c2 = 0..7 of one key row are ascending), in the
    // log2 domain, 4 copies shifted by 0..3 elements so that every 8-key run starts 16-byte aligned
    for (int e = tid; e < 4 * 4 * Cfg::BIAS_STRIDE; e += 128) {
      const int sh = e / (4 * Cfg::BIAS_STRIDE), rem = e - sh * 4 * Cfg::BIAS_STRIDE;
      const int h = rem / Cfg::BIAS_STRIDE, idx = rem - h * Cfg::BIAS_STRIDE + sh;
      sBias[e] = idx <= 224 ? __ldg(p.bias_table + (224 - idx) * kHeads + hgc * kHG + h) * 1.4426950408889634f : 0.f;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  if (tid == 0) TC_TS(0, 0);
  if (tid == 128) TC_TS(1, 0);

  uint32_t U = 0;        // units (source, head pair) processed so far
  uint32_t J = 0;        // active sources processed so far
  float m_run[2] = {-INFINITY, -INFINITY};

  for (int j0 = 0; j0 < nrec; j0 += Cfg::kMaxSrc) {
    const int nsrc = min(Cfg::kMaxSrc, nrec - j0);
    // ---- taps + visibility of every (source, token) of this group ----
    if (j0 != 0) {
      __syncthreads();
      if (tid < nsrc) {
        sMap[tid].valid = p.cav_mask[b * p.L + j0 + tid] != 0 ? 1 : 0;
        sMap[tid].wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + j0 + tid) * p.L + i) * 16, p.H, p.W, p.cell);
      }
      stage_kv_bias(p, sKvb, b, j0, nsrc, te, hgc, tid, Cfg::THREADS);
    }
    if (tid < Cfg::kMaxSrc) sAnyVis[tid] = 0;
    __syncthreads();
    for (int e = tid; e < nsrc * kS; e += Cfg::THREADS) {
      const int js = e >> 6, tk = e & 63, j = j0 + js;
      TapRec rec; rec.x0 = 0; rec.y0 = 0; rec.w01 = 0; rec.w23 = 0;
      if (sMap[js].valid) {
        int r, c; group_token(p.kind, gy, gx, tk, p.H, p.W, r, c);
        double sx, sy; warp_src(sMap[js].wm, c, r, sx, sy);
        bool vis = warp_visible(sx, sy, p.H, p.W);
        if (p.key_mask != nullptr && p.key_mask[static_cast<size_t>(b * p.L + j) * N + r * p.W + c] == 0) vis = false;
        if (vis) {
          const Taps tp = make_taps(sx, sy, p.H, p.W);
          rec.x0 = static_cast<short>(tp.x0); rec.y0 = static_cast<short>(tp.y0);
          rec.w01 = pack_bf16x2(tp.w00, tp.w01); rec.w23 = pack_bf16x2(tp.w10, tp.w11);
          atomicOr(&sAnyVis[js], 1);
#if (HMVIT_TC_DBG & 4)   // measured: no gain (tools/micro/gather_bw.cu, profiles/r1_attention_study.md)
          // pull the (up to) 4 tap rows of K' and V' into L2 now: the gather warps' loads then see L2 latency
          // instead of HBM latency (a warp keeps only ~2 KB of loads in flight, so latency is throughput)
          const size_t plane_b = static_cast<size_t>(p.B) * p.L * N * 512;           // bytes per te plane
          const size_t row0 = (static_cast<size_t>(b * p.L + j) * N) * 512 + static_cast<size_t>(te) * plane_b + cu0 * 16;
          const float wt[4] = {tp.w00, tp.w01, tp.w10, tp.w11};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (wt[q] != 0.f) {
              const size_t off = row0 + static_cast<size_t>((tp.y0 + (q >> 1)) * p.W + tp.x0 + (q & 1)) * 512;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(reinterpret_cast<const uint8_t*>(p.k) + off) : "memory");
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(reinterpret_cast<const uint8_t*>(p.v) + off) : "memory");
            }
          }
#endif
        }
      }
      sTapAll[e] = rec;
    }
    __syncthreads();
    if (tid == 0) {
      int n = 0;
      for (int js = 0; js < nsrc; ++js) if (sAnyVis[js] != 0) sAct[n++] = js;
      sAct[Cfg::kMaxSrc] = n;
    }
    __syncthreads();
    const int nact = sAct[Cfg::kMaxSrc];

    if (warp < 4) {
      // =========================================== SOFTMAX ===========================================
      const int hh = tid >> 6, row = tid & 63;
      const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
      const int E = 112 - (row & 7) - 15 * (row >> 3);
      for (int ja = 0; ja < nact; ++ja) {
        const bool first_src = (J == 0);
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          if (tid == 0) TC_TS(0, 1 + U * 4 + 0);
          mbar_wait(s_full, U & 1u);
          tc_fence_after();
          if (tid == 0) TC_TS(0, 1 + U * 4 + 1);
          // the 64 logits of this row are processed as two halves of 32 keys; the first half (bias added) is
          // parked in its own TMEM columns while the second one is reduced, so that at most 32 logits are live
          // in registers (96-register budget at 2 CTAs / SM)
          const float* bt = sBias + (pr * 2 + hh) * Cfg::BIAS_STRIDE;
          float mx = -INFINITY;
          auto add_bias = [&](uint32_t (&sv)[32], int r2base) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const int e0 = E + 15 * (r2base + rr), sh = e0 & 3;
              const float4* bp = reinterpret_cast<const float4*>(bt + sh * (4 * Cfg::BIAS_STRIDE) + (e0 - sh));
              const float4 b0 = bp[0], b1 = bp[1];
              uint32_t* s8 = sv + rr * 8;
              float v;
              v = __uint_as_float(s8[0]) + b0.x; s8[0] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[1]) + b0.y; s8[1] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[2]) + b0.z; s8[2] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[3]) + b0.w; s8[3] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[4]) + b1.x; s8[4] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[5]) + b1.y; s8[5] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[6]) + b1.z; s8[6] = __float_as_uint(v); mx = fmaxf(mx, v);
              v = __uint_as_float(s8[7]) + b1.w; s8[7] = __float_as_uint(v); mx = fmaxf(mx, v);
            }
          };
          uint32_t pk[32];
          float m_new, alpha;
          {
            uint32_t sa[32];
            tmem_ld32(tm + lane_base + Cfg::TM_S, sa);
            tmem_ld_wait();
            if (!(HMVIT_TC_DBG & 2)) add_bias(sa, 0);
            tmem_st32(tm + lane_base + Cfg::TM_S, sa);
          }
          {
            uint32_t sb[32];
            tmem_ld32(tm + lane_base + Cfg::TM_S + 32, sb);
            tmem_ld_wait();
            if (!(HMVIT_TC_DBG & 2)) add_bias(sb, 4);
            else mx = 0.f;
            m_new = fmaxf(m_run[pr], mx);
            alpha = ex2(m_run[pr] - m_new);                    // 0 for the first source (m_run = -inf)
            m_run[pr] = m_new;
#pragma unroll
            for (int k = 0; k < 16; ++k)
              pk[16 + k] = (HMVIT_TC_DBG & 2) ? sb[k] : pack_bf16x2(ex2(__uint_as_float(sb[2 * k]) - m_new), ex2(__uint_as_float(sb[2 * k + 1]) - m_new));
          }
          {
            uint32_t sa[32];
            tmem_st_wait();
            tmem_ld32(tm + lane_base + Cfg::TM_S, sa);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free);                    // S is consumed: the next QK^T may overwrite it
#pragma unroll
            for (int k = 0; k < 16; ++k)
              pk[k] = (HMVIT_TC_DBG & 2) ? sa[k] : pack_bf16x2(ex2(__uint_as_float(sa[2 * k]) - m_new), ex2(__uint_as_float(sa[2 * k + 1]) - m_new));
          }
          if (tid == 0) TC_TS(0, 1 + U * 4 + 2);
          if (U > 0) { mbar_wait(p_empty, (U - 1) & 1u); tc_fence_after(); }   // P tile free; PV of the previous source retired
          if (tid == 0) TC_TS(0, 1 + U * 4 + 3);
          if (!first_src && !__all_sync(0xffffffffu, alpha == 1.0f)) {
            // the running max moved: rescale this pair's accumulators in TMEM
            const uint32_t dcol = tm + lane_base + (pr ? Cfg::TM_D1 : Cfg::TM_D0) + hh * 32;
            const uint32_t lcol = tm + lane_base + (pr ? Cfg::TM_L1 : Cfg::TM_L0);
            uint32_t d[32], l1;
            tmem_ld32(dcol, d);
            tmem_ld1(lcol, l1);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) d[k] = __float_as_uint(__uint_as_float(d[k]) * alpha);
            l1 = __float_as_uint(__uint_as_float(l1) * alpha);
            tmem_st32(dcol, d);
            tmem_st1(lcol, l1);
          }
          tmem_st32(tm + lane_base + Cfg::TM_P, pk);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(p_full);
          ++U;
        }
        ++J;
      }
    } else if (warp < 8) {
      // =========================================== GATHER ===========================================
      // Flat sequence of work items (active source, K | V, batch of 2 tokens per half-warp): the 8 loads of
      // item w + 1 are issued before item w is blended, so tap loads are in flight all the time.
      const int gw = warp - 4, hl = lane >> 4, u16 = lane & 15;
      const int pr = u16 >> 3, un = u16 & 7;
      const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;            // uint4 units per te plane
      const uint4* kbase = reinterpret_cast<const uint4*>(p.k) + te * plane + cu0 + u16;
      const uint4* vbase = reinterpret_cast<const uint4*>(p.v) + te * plane + cu0 + u16;
      const int items = nact * 8;
      auto issue = [&](int w, uint4 (&tv)[2][4]) {
        const int js = sAct[w >> 3], kv = (w >> 2) & 1, n = w & 3;
        const uint4* src = (kv == 0 ? kbase : vbase) + static_cast<size_t>(b * p.L + j0 + js) * N * 32;
        const TapRec* sTap = sTapAll + js * kS;
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const TapRec rec = sTap[gw * 16 + (n * 2 + t2) * 2 + hl];
          const uint32_t wq[4] = {rec.w01 & 0xffffu, rec.w01 >> 16, rec.w23 & 0xffffu, rec.w23 >> 16};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            tv[t2][q] = make_uint4(0, 0, 0, 0);
            if (wq[q] != 0u && !(HMVIT_TC_DBG & 1)) {
              const int yy = rec.y0 + (q >> 1), xx = rec.x0 + (q & 1);
              tv[t2][q] = __ldg(src + static_cast<size_t>(yy * p.W + xx) * 32);
            }
          }
        }
      };
      auto process = [&](int w, uint4 (&tv)[2][4]) {
        const int js = sAct[w >> 3], kv = (w >> 2) & 1, n = w & 3;
        const uint32_t Jw = J + (w >> 3), vst = Jw & 1u;
        const TapRec* sTap = sTapAll + js * kS;
        if (n == 0) {                                         // first store into this tile: wait until its readers retired
          if (tid == 128) TC_TS(1, 1 + Jw * 6 + 1 + kv * 3);
          if (kv == 0) mbar_wait(k_empty, (Jw & 1u) ^ 1u);
          else mbar_wait(&v_empty[vst], ((Jw >> 1) & 1u) ^ 1u);
          if (tid == 128) TC_TS(1, 1 + Jw * 6 + 2 + kv * 3);
        }
        uint8_t* dstT = (kv == 0 ? sK : sV + vst * 16384) + pr * 8192;
        const uint4 bb = *reinterpret_cast<const uint4*>(sKvb + (js * 2 + kv) * 256 + u16 * 16);
#pragma unroll
        for (int t2 = 0; t2 < 2; ++t2) {
          const int s = gw * 16 + (n * 2 + t2) * 2 + hl;
          const TapRec rec = sTap[s];
          uint4 o = make_uint4(0, 0, 0, 0);
          if ((rec.w01 | rec.w23) != 0u) {
            const uint32_t wq[4] = {rec.w01 & 0xffffu, rec.w01 >> 16, rec.w23 & 0xffffu, rec.w23 >> 16};
            o = bb;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t w2 = wq[q] | (wq[q] << 16);
              o.x = hfma2_bf16(w2, tv[t2][q].x, o.x); o.y = hfma2_bf16(w2, tv[t2][q].y, o.y);
              o.z = hfma2_bf16(w2, tv[t2][q].z, o.z); o.w = hfma2_bf16(w2, tv[t2][q].w, o.w);
            }
          }
          *reinterpret_cast<uint4*>(dstT + sw128_offset(s, un)) = o;
        }
        if (n == 3) {
          if (kv == 1 && tid - 128 < 8) {
            // visibility row of this source: 8 keys per 16-byte unit, bf16 1.0 / 0.0
            const int u = tid - 128;
            uint32_t wv[4];
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              const TapRec r0 = sTap[u * 8 + k2 * 2], r1 = sTap[u * 8 + k2 * 2 + 1];
              wv[k2] = ((r0.w01 | r0.w23) != 0u ? 0x3F80u : 0u) | ((r1.w01 | r1.w23) != 0u ? 0x3F800000u : 0u);
            }
            *reinterpret_cast<uint4*>(sVis + vst * 2048 + u * 16) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          }
          fence_proxy_async_smem();
          mbar_arrive(kv == 0 ? k_full : &v_full[vst]);
          if (tid == 128) TC_TS(1, 1 + Jw * 6 + 3 + kv * 3);
        }
      };
      if (items > 0) {
        uint4 ta[2][4], tb[2][4];
        if (tid == 128) TC_TS(1, 1 + J * 6 + 0);
        issue(0, ta);
#pragma unroll 1
        for (int w = 0; w < items; w += 2) {
          issue(w + 1, tb);                                   // items is even
          process(w, ta);
          if (w + 2 < items) issue(w + 2, ta);
          process(w + 1, tb);
        }
      }
      J += nact;
    } else {
      // =========================================== MMA ===========================================
      if (lane == 0 && nact > 0) {
        constexpr uint32_t idesc_qk = umma_idesc(1u, 128, 64);
        constexpr uint32_t idesc_pv = umma_idesc(1u, 128, 64) | (1u << 16);      // B (V tile) MN-major
        constexpr uint32_t idesc_l = umma_idesc(1u, 128, 16);
        const uint32_t q_u = smem_u32(sQ), k_u = smem_u32(sK), v_u = smem_u32(sV), vis_u = smem_u32(sVis);
        auto issue_qk = [&](uint32_t u) {                // S = Qbd_pr K_pr^T
          const uint32_t pr = u & 1u;
          mbar_wait(s_free, (u & 1u) ^ 1u);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss<2>(tm + Cfg::TM_S, umma_desc_sw128(q_u + pr * 16384 + ks * 32), umma_desc_sw128(k_u + pr * 8192 + ks * 32),
                       idesc_qk, ks != 0 ? 1u : 0u);
          umma_commit(s_full);
        };
        auto issue_pv = [&](uint32_t u, uint32_t vst, bool first) {    // D_pr (+)= P V_pr ; Lsum_pr (+)= P vis
          const uint32_t pr = u & 1u;
          mbar_wait(p_full, u & 1u);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t pa = tm + Cfg::TM_P + ks * 8;
            umma_ts_bf16(tm + (pr ? Cfg::TM_D1 : Cfg::TM_D0), pa, umma_desc_sw128_mn(v_u + vst * 16384 + pr * 8192 + ks * 2048),
                         idesc_pv, (!first || ks != 0) ? 1u : 0u);
            umma_ts_bf16(tm + (pr ? Cfg::TM_L1 : Cfg::TM_L0), pa, umma_desc_sw128(vis_u + vst * 2048 + ks * 32), idesc_l,
                         (!first || ks != 0) ? 1u : 0u);
          }
          umma_commit(p_empty);
        };
        mbar_wait(k_full, J & 1u);
        tc_fence_after();
        issue_qk(U);
        for (int jj = 0; jj < nact; ++jj) {
          const bool first = (J == 0);
          issue_qk(U + 1);
          umma_commit(k_empty);                           // both pairs' QK^T of this source issued
          const uint32_t vst = J & 1u;
          mbar_wait(&v_full[vst], (J >> 1) & 1u);
          tc_fence_after();
          issue_pv(U, vst, first);
          if (jj + 1 < nact) {
            mbar_wait(k_full, (J + 1) & 1u);
            tc_fence_after();
            issue_qk(U + 2);
          }
          issue_pv(U + 1, vst, first);
          umma_commit(&v_empty[vst]);
          U += 2; ++J;
        }
      } else {
        U += 2 * nact; J += nact;
      }
      __syncwarp();
    }
  }

  // ------------------------------ normalise and store ------------------------------
  if (warp < 4) {
    const int hh = tid >> 6, row = tid & 63;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    int r, c; group_token(p.kind, gy, gx, row, p.H, p.W, r, c);
    if (U > 0) { mbar_wait(p_empty, (U - 1) & 1u); tc_fence_after(); }
    if (tid == 0) TC_TS(0, 1 + U * 4);
#pragma unroll 1
    for (int pr = 0; pr < 2; ++pr) {
      uint32_t d[32], l1 = 0;
      if (U > 0) {
        tmem_ld32(tm + lane_base + (pr ? Cfg::TM_D1 : Cfg::TM_D0) + hh * 32, d);
        tmem_ld1(tm + lane_base + (pr ? Cfg::TM_L1 : Cfg::TM_L0), l1);
        tmem_ld_wait();
      }
      const float l = __uint_as_float(l1);
      const float il = (U > 0 && l > 0.f) ? 1.0f / l : 0.f;
      uint4* dst = reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(a) * N + r * p.W + c) * kC + (hgc * kHG + pr * 2 + hh) * kDh);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint4 o;
        if (il == 0.f) o = make_uint4(0, 0, 0, 0);
        else {
          o.x = pack_bf16x2(__uint_as_float(d[u * 8 + 0]) * il, __uint_as_float(d[u * 8 + 1]) * il);
          o.y = pack_bf16x2(__uint_as_float(d[u * 8 + 2]) * il, __uint_as_float(d[u * 8 + 3]) * il);
          o.z = pack_bf16x2(__uint_as_float(d[u * 8 + 4]) * il, __uint_as_float(d[u * 8 + 5]) * il);
          o.w = pack_bf16x2(__uint_as_float(d[u * 8 + 6]) * il, __uint_as_float(d[u * 8 + 7]) * il);
        }
        dst[u] = o;
      }
    }
  }
  if (tid == 0) TC_TS(0, 2 + U * 4);
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TM_COLS>(tm);
  }
}

}  // namespace hmvit
