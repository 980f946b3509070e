// Backward of the fused warp + mask + multi-agent group attention (attn.cuh) for one
// (scene b, ego agent i, token group g, head group of 4 heads).
//
// Adjoint of the reference's warp_features + HeteroAttention.forward core
// (hetero_fusion.py:338-361, 187-277; autograd through F.grid_sample / einsum / softmax there).
// Per source agent j the kernel re-gathers the projected keys / values exactly like the forward
// (same fp64 source coordinates, same bf16 tap blend), recomputes the probabilities from the saved
// log-sum-exp, and produces
//   dQ'            (ego rows, fp32 atomics -- two warps own disjoint key halves of the same head)
//   dK'|te, dV'|te (bilinear SCATTER of dKg / dVg through the 4 taps into the source agent's rows, fp32 atomics)
//   d bk, d bv     (folded key / value biases, per (ego type, source type))
//   d bias_table   (relative position bias; accumulated in shared memory, flushed once per CTA)
// With  S2 = Q' Kg^T + bias log2(e)  (log2 domain),  P = 2^(S2 - lse2),  D = rowsum(dO o O):
//   dVg = P^T dO ;  dP = dO Vg^T ;  dSn = P o (dP - D) ;  d bias_table += dSn ;
//   dS2 = ln(2) dSn ;  dQ' = dS2 Kg ;  dKg = dS2^T Q'.
// Contractions use warp-level wmma tiles (bf16 operands, fp32 accumulate): warp w owns head (w & 3) of
// the head group and the key half (w >> 2) of every source.  Round-1 kernel: correctness first.
#pragma once
#include "attn.cuh"
#include "wmma_shared.cuh"

#ifndef HMVIT_BWD_DBG   // bottleneck-hunting builds only (results are wrong): 1 no shared bias-gradient atomics, 2 no global scatter,
                        // 4 no per-step warp barrier in the probability loop, 8 no probability loop, 16 no K/V re-gather, 32 no tensor-core tiles
#define HMVIT_BWD_DBG 0
#endif

namespace hmvit {

// 128-bit vector reduction into global memory (sm_90+): one instruction adds four consecutive fp32 values
HMVIT_DEVINL void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct AttnBwdParams {
  int B, L, H, W;
  int kind, ego_only;
  const int* mode; const int* record_len; const int* cav_mask;
  const float* T; double cell;
  const __nv_bfloat16* q;      // [R][256]
  const __nv_bfloat16* k;      // [2][R][256]
  const __nv_bfloat16* v;      // [2][R][256]
  const float* bk; const float* bv; const float* bias_table;
  const __nv_bfloat16* o;      // forward output [R][256]
  const __nv_bfloat16* d_o;    // gradient w.r.t. the forward output [R][256]
  const float* lse;            // [R][8] log2-domain log-sum-exp saved by the forward
  float* dq;                   // [R][256] fp32, accumulated (caller zero-fills)
  float* dk;                   // [2][R][256]
  float* dv;                   // [2][R][256]
  float* dbk; float* dbv;      // [2][2][256], accumulated
  float* dbias_table;          // [225][8], accumulated
};

struct AttnBwdCfg {
  static constexpr int THREADS = 256;
  static constexpr int LD = 136;                                  // bf16 elements per tile row (128 + 8 pad)
  static constexpr int TILE_BYTES = kS * LD * 2;                  // 17408
  static constexpr int SCR_BYTES = 4096 + 4096 + 2048 + 2048;     // per warp: S | dP (fp32 [64][16]) | P | dS (bf16 [64][16])
  static constexpr int OFF_SCR = 6 * TILE_BYTES;                  // Q | dO | (Kg | Vg) x 2: the re-gathered tiles are double-buffered
  static constexpr int OFF_BIAS = OFF_SCR + 8 * SCR_BYTES;        // [4][232] fp32, log2 domain
  static constexpr int OFF_BGRAD = OFF_BIAS + kHG * kBiasStride * 4;
  static constexpr int OFF_D = OFF_BGRAD + 8 * kBiasStride * 4;   // bias gradient: one private [232] table per warp; D: [4][64]
  static constexpr int OFF_LSE = OFF_D + kHG * kS * 4;            // [4][64]
  static constexpr int OFF_TAP = OFF_LSE + kHG * kS * 4;
  static constexpr int OFF_VIS = OFF_TAP + kMaxSrc * kS * static_cast<int>(sizeof(TapRec));
  static constexpr int SMEM_BYTES = OFF_VIS + 32;
};

__global__ void __launch_bounds__(AttnBwdCfg::THREADS, 1) group_attn_bwd_kernel(const AttnBwdParams p) {
  using namespace nvcuda;
  using Cfg = AttnBwdCfg;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);            // a malformed record_len must not index past the scene's slots
  if (i >= nrec || (p.ego_only && i != 0)) return;
  const int N = p.H * p.W;
  const int GX = p.W / kWin;
  const int grp = blockIdx.x >> 1, hgc = blockIdx.x & 1;
  const int gy = grp / GX, gx = grp - gy * GX;
  const int te = p.mode[a] != 0 ? 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hl = warp & 3, kh = warp >> 2;                      // head within the group, key half
  const int hlf = lane >> 4, u16 = lane & 15;                   // gather: half-warp per token, 16-byte unit
  const size_t R = static_cast<size_t>(p.B) * p.L * N;

  extern __shared__ __align__(128) uint8_t smem[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* sdO = reinterpret_cast<__nv_bfloat16*>(smem + Cfg::TILE_BYTES);
  __nv_bfloat16* sKV = reinterpret_cast<__nv_bfloat16*>(smem + 2 * Cfg::TILE_BYTES);      // [2 buffers][Kg | Vg]
  uint8_t* scr = smem + Cfg::OFF_SCR + warp * Cfg::SCR_BYTES;
  float* sS = reinterpret_cast<float*>(scr);                    // [64][16] logits, later dKg chunk [16][32]
  float* sdP = reinterpret_cast<float*>(scr + 4096);            // [64][16] dP, later dVg chunk [16][32]
  __nv_bfloat16* sPb = reinterpret_cast<__nv_bfloat16*>(scr + 8192);
  __nv_bfloat16* sdSb = reinterpret_cast<__nv_bfloat16*>(scr + 8192 + 2048);
  float* sBias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  float* sBgrad = reinterpret_cast<float*>(smem + Cfg::OFF_BGRAD);
  float* sD = reinterpret_cast<float*>(smem + Cfg::OFF_D);
  float* sLse = reinterpret_cast<float*>(smem + Cfg::OFF_LSE);
  TapRec* sTapAll = reinterpret_cast<TapRec*>(smem + Cfg::OFF_TAP);
  int* sAnyVis = reinterpret_cast<int*>(smem + Cfg::OFF_VIS);

  const int cu0 = hgc * 16;                                     // first 16-byte unit of the head group in a 512-byte row
  // ---- stage Q, dO; D = rowsum_head(dO o O); lse; bias table ----
  {
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    const uint4* gsrc = reinterpret_cast<const uint4*>(p.d_o) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    const uint4* osrc = reinterpret_cast<const uint4*>(p.o) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    for (int tt = 0; tt < 4; ++tt) {
      const int s = tt * 16 + warp * 2 + hlf;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      const size_t off = static_cast<size_t>(r * p.W + c) * 32;
      const uint4 qv = __ldg(qsrc + off), gv = __ldg(gsrc + off), ov = __ldg(osrc + off);
      *reinterpret_cast<uint4*>(sQ + s * Cfg::LD + u16 * 8) = qv;
      *reinterpret_cast<uint4*>(sdO + s * Cfg::LD + u16 * 8) = gv;
      float d = bf16_lo(gv.x) * bf16_lo(ov.x) + bf16_hi(gv.x) * bf16_hi(ov.x) + bf16_lo(gv.y) * bf16_lo(ov.y) + bf16_hi(gv.y) * bf16_hi(ov.y) +
                bf16_lo(gv.z) * bf16_lo(ov.z) + bf16_hi(gv.z) * bf16_hi(ov.z) + bf16_lo(gv.w) * bf16_lo(ov.w) + bf16_hi(gv.w) * bf16_hi(ov.w);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);                  // 4 lanes (32 channels) = one head
      if ((u16 & 3) == 0) {
        const int h = u16 >> 2;
        sD[h * kS + s] = d;
        sLse[h * kS + s] = __ldg(p.lse + (static_cast<size_t>(a) * N + r * p.W + c) * kHeads + hgc * kHG + h);
      }
    }
    for (int e = threadIdx.x; e < 225 * kHG; e += Cfg::THREADS) {
      const int idx = e >> 2, h = e & 3;
      sBias[h * kBiasStride + idx] = __ldg(p.bias_table + idx * kHeads + hgc * kHG + h) * 1.4426950408889634f;
      sBgrad[h * kBiasStride + idx] = 0.f;
      sBgrad[(h + 4) * kBiasStride + idx] = 0.f;
    }
  }

  wmma::fragment<wmma::accumulator, 16, 16, 16, float> dq[4][2];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) wmma::fill_fragment(dq[mt][nt], 0.f);

  const size_t plane = R * 32;                                   // uint4 units per te plane
  const int ch = hgc * 128 + hl * 32 + lane;                     // channel this lane scatters

  int nvis_done = 0;                                             // sources processed so far (CTA-uniform): tile buffer parity
  for (int j0 = 0; j0 < nrec; j0 += kMaxSrc) {
    const int nsrc = min(kMaxSrc, nrec - j0);
    __syncthreads();
    if (threadIdx.x < kMaxSrc) sAnyVis[threadIdx.x] = 0;
    __syncthreads();
    // ---- taps + visibility, identical to the forward ----
    for (int e = threadIdx.x; e < nsrc * kS; e += Cfg::THREADS) {
      const int js = e >> 6, tk = e & 63, j = j0 + js;
      TapRec rec; rec.x0 = 0; rec.y0 = 0; rec.w01 = 0; rec.w23 = 0;
      if (p.cav_mask[b * p.L + j] != 0) {
        const WarpMap wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + j) * p.L + i) * 16, p.H, p.W, p.cell);
        int r, c; group_token(p.kind, gy, gx, tk, p.H, p.W, r, c);
        double sx, sy; warp_src(wm, c, r, sx, sy);
        if (warp_visible(sx, sy, p.H, p.W)) {
          const Taps tp = make_taps(sx, sy, p.H, p.W);
          rec.x0 = static_cast<short>(tp.x0); rec.y0 = static_cast<short>(tp.y0);
          rec.w01 = pack_bf16x2(tp.w00, tp.w01); rec.w23 = pack_bf16x2(tp.w10, tp.w11);
          atomicOr(&sAnyVis[js], 1);
        }
      }
      sTapAll[e] = rec;
    }
    __syncthreads();

    for (int js = 0; js < nsrc; ++js) {
      const int j = j0 + js;
      if (sAnyVis[js] == 0) continue;
      // Two tile buffers: a warp that has finished source j gathers source j + 1 into the other buffer while slower warps
      // still multiply source j -- one CTA barrier per source (tile complete) instead of two.  The buffer written now was
      // last read for the source before the previous one, and every warp passed the previous source's barrier after that.
      __nv_bfloat16* sK = sKV + (nvis_done & 1) * (2 * Cfg::TILE_BYTES / 2);
      __nv_bfloat16* sV = sK + Cfg::TILE_BYTES / 2;
      ++nvis_done;
      const TapRec* sTap = sTapAll + js * kS;
      const int tj = p.mode[b * p.L + j] != 0 ? 1 : 0;
      // ---- re-gather Kg / Vg of source j (same bf16 blend as the forward) ----
      {
        const uint4* ksrc = reinterpret_cast<const uint4*>(p.k) + te * plane + static_cast<size_t>(b * p.L + j) * N * 32 + cu0 + u16;
        const uint4* vsrc = reinterpret_cast<const uint4*>(p.v) + te * plane + static_cast<size_t>(b * p.L + j) * N * 32 + cu0 + u16;
        uint32_t bk2[4], bv2[4];
        {
          const float4* pk = reinterpret_cast<const float4*>(p.bk + (te * 2 + tj) * kC + (cu0 + u16) * 8);
          const float4* pv = reinterpret_cast<const float4*>(p.bv + (te * 2 + tj) * kC + (cu0 + u16) * 8);
          const float4 k0 = __ldg(pk), k1 = __ldg(pk + 1), v0 = __ldg(pv), v1 = __ldg(pv + 1);
          bk2[0] = pack_bf16x2(k0.x, k0.y); bk2[1] = pack_bf16x2(k0.z, k0.w); bk2[2] = pack_bf16x2(k1.x, k1.y); bk2[3] = pack_bf16x2(k1.z, k1.w);
          bv2[0] = pack_bf16x2(v0.x, v0.y); bv2[1] = pack_bf16x2(v0.z, v0.w); bv2[2] = pack_bf16x2(v1.x, v1.y); bv2[3] = pack_bf16x2(v1.z, v1.w);
        }
        // The tap loads of token tt + 1 are issued before token tt is blended (two register images of the eight 16-byte
        // taps): the four tokens of a half-warp exposed four global round trips per source, 17 % of the kernel.
        auto load_taps = [&](const TapRec& rec, uint4 (&kk)[4], uint4 (&vv)[4]) {
          const uint32_t wq[4] = {rec.w01 & 0xffffu, rec.w01 >> 16, rec.w23 & 0xffffu, rec.w23 >> 16};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            kk[q] = make_uint4(0, 0, 0, 0); vv[q] = make_uint4(0, 0, 0, 0);
            if (!(HMVIT_BWD_DBG & 16) && wq[q] != 0u) {
              const int yy = rec.y0 + (q >> 1), xx = rec.x0 + (q & 1);
              const size_t off = static_cast<size_t>(yy * p.W + xx) * 32;
              kk[q] = __ldg(ksrc + off); vv[q] = __ldg(vsrc + off);
            }
          }
        };
        auto blend_store = [&](const TapRec& rec, const uint4 (&kk)[4], const uint4 (&vv)[4], int s) {
          uint4 ko = make_uint4(0, 0, 0, 0), vo = make_uint4(0, 0, 0, 0);
          if (!(HMVIT_BWD_DBG & 16) && (rec.w01 | rec.w23) != 0u) {
            const uint32_t wq[4] = {rec.w01 & 0xffffu, rec.w01 >> 16, rec.w23 & 0xffffu, rec.w23 >> 16};
            ko = make_uint4(bk2[0], bk2[1], bk2[2], bk2[3]);
            vo = make_uint4(bv2[0], bv2[1], bv2[2], bv2[3]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {                      // same taps, order and arithmetic as the forward
              if (wq[q] == 0u) continue;
              const uint32_t w2 = wq[q] | (wq[q] << 16);
              ko.x = hfma2_bf16(w2, kk[q].x, ko.x); ko.y = hfma2_bf16(w2, kk[q].y, ko.y);
              ko.z = hfma2_bf16(w2, kk[q].z, ko.z); ko.w = hfma2_bf16(w2, kk[q].w, ko.w);
              vo.x = hfma2_bf16(w2, vv[q].x, vo.x); vo.y = hfma2_bf16(w2, vv[q].y, vo.y);
              vo.z = hfma2_bf16(w2, vv[q].z, vo.z); vo.w = hfma2_bf16(w2, vv[q].w, vo.w);
            }
          }
          *reinterpret_cast<uint4*>(sK + s * Cfg::LD + u16 * 8) = ko;
          *reinterpret_cast<uint4*>(sV + s * Cfg::LD + u16 * 8) = vo;
        };
        TapRec recs[4];
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) recs[tt] = sTap[tt * 16 + warp * 2 + hlf];
        uint4 kA[4], vA[4], kB[4], vB[4];
        load_taps(recs[0], kA, vA);
        load_taps(recs[1], kB, vB);
        blend_store(recs[0], kA, vA, 0 * 16 + warp * 2 + hlf);
        load_taps(recs[2], kA, vA);
        blend_store(recs[1], kB, vB, 1 * 16 + warp * 2 + hlf);
        load_taps(recs[3], kB, vB);
        blend_store(recs[2], kA, vA, 2 * 16 + warp * 2 + hlf);
        blend_store(recs[3], kB, vB, 3 * 16 + warp * 2 + hlf);
      }
      __syncthreads();

      float4 bsum_k4 = make_float4(0.f, 0.f, 0.f, 0.f), bsum_v4 = make_float4(0.f, 0.f, 0.f, 0.f);   // lane: 4 channels, keys = lane / 8 (mod 4)
      for (int kc = kh * 2; kc < kh * 2 + 2; ++kc) {
        // a chunk of 16 keys without a visible key contributes nothing (P = 0): skip it (exact; ~1/3 of the chunks of
        // the collaborators' sources at config 2)
        {
          const TapRec rv = sTap[kc * 16 + (lane & 15)];
          if (__ballot_sync(0xffffffffu, (rv.w01 | rv.w23) != 0u) == 0u) continue;
        }
        // ---- S = Q_h Kg_h^T, dP = dO_h Vg_h^T for 64 queries x 16 keys ----
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          wmma::fragment<wmma::accumulator, 16, 16, 16, float> sacc, pacc;
          wmma::fill_fragment(sacc, 0.f); wmma::fill_fragment(pacc, 0.f);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fq, fg;
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> fk, fv;
            wmma_load_shared(fq, sQ + (mt * 16) * Cfg::LD + hl * 32 + ks * 16, Cfg::LD);
            wmma_load_shared(fg, sdO + (mt * 16) * Cfg::LD + hl * 32 + ks * 16, Cfg::LD);
            wmma_load_shared(fk, sK + (kc * 16) * Cfg::LD + hl * 32 + ks * 16, Cfg::LD);
            wmma_load_shared(fv, sV + (kc * 16) * Cfg::LD + hl * 32 + ks * 16, Cfg::LD);
            wmma::mma_sync(sacc, fq, fk, sacc);
            wmma::mma_sync(pacc, fg, fv, pacc);
          }
          wmma::store_matrix_sync(sS + mt * 256, sacc, 16, wmma::mem_row_major);
          wmma::store_matrix_sync(sdP + mt * 256, pacc, 16, wmma::mem_row_major);
        }
        __syncwarp();
        // ---- probabilities and logit gradients ----
        // The relative position bias gradient goes into a table PRIVATE to this warp with plain read-modify-writes
        // (shared-memory float atomics cost 22 % of the kernel).  This loop was 32 % of the kernel and issue-bound (two warps
        // per scheduler, ~40 instructions per logit), so a lane takes TWO adjacent keys of the chunk (one 64-bit load of S / dP,
        // one packed bf16x2 store of P / dS, the query's lse / D and the index arithmetic shared) for one query per step:
        // lane = (q, key pair kp), query row 2 q + (e >> 3), column (e & 7) ^ (q & 1).  The four queries of a step lie in
        // different window rows (2 q - key row is distinct over (q, key row)), so the 64 (query, key) pairs of a step hit 64
        // different table entries; the column flip of odd q keeps the 64-bit scratch reads of a half-warp on different banks.
        float* sBgW = sBgrad + warp * kBiasStride;
        {
          const int kp = lane & 7, qs = lane >> 3;
          const int kk0 = kp * 2, sp0 = kc * 16 + kk0;            // keys sp0, sp0 + 1 (same key row)
          const TapRec rk0 = sTap[sp0], rk1 = sTap[sp0 + 1];
          const bool v0 = (rk0.w01 | rk0.w23) != 0u, v1 = (rk1.w01 | rk1.w23) != 0u;
          const int ky = sp0 >> 3, kx0 = sp0 & 7;
          const float* sBiasH = sBias + hl * kBiasStride;
#pragma unroll 4
          for (int e = 0; e < ((HMVIT_BWD_DBG & 8) ? 0 : 16); ++e) {
            const int sy = 2 * qs + (e >> 3), sx = (e & 7) ^ (qs & 1);
            const int s = sy * 8 + sx, idx = s * 16 + kk0;
            const int rel = (sy - ky + 7) * 15 + (sx - kx0 + 7);  // of key sp0; key sp0 + 1: rel - 1 (>= 0: kx0 + 1 <= 7)
            const float2 sv = *reinterpret_cast<const float2*>(sS + idx);
            const float2 dp = *reinterpret_cast<const float2*>(sdP + idx);
            const float l2 = sLse[hl * kS + s], dd = sD[hl * kS + s];
            float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
            if (v0) { p0 = ex2(sv.x + sBiasH[rel] - l2); d0 = p0 * (dp.x - dd); }
            if (v1) { p1 = ex2(sv.y + sBiasH[rel - 1] - l2); d1 = p1 * (dp.y - dd); }
#if !(HMVIT_BWD_DBG & 1)
            if (v0) sBgW[rel] += d0;
            if (v1) sBgW[rel - 1] += d1;
#endif
            *reinterpret_cast<uint32_t*>(sPb + idx) = pack_bf16x2(p0, p1);
            *reinterpret_cast<uint32_t*>(sdSb + idx) = pack_bf16x2(d0 * 0.69314718055994530942f, d1 * 0.69314718055994530942f);
#if !(HMVIT_BWD_DBG & 4)
            __syncwarp();                                        // the next step may touch the same table entry from another lane
#endif
          }
        }
        __syncwarp();
        // ---- dQ_h += dS Kg_h ----
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
          wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fs;
          wmma_load_shared(fs, sdSb + mt * 256, 16);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fk;
            wmma_load_shared(fk, sK + (kc * 16) * Cfg::LD + hl * 32 + nt * 16, Cfg::LD);
            wmma::mma_sync(dq[mt][nt], fs, fk, dq[mt][nt]);
          }
        }
        // ---- dKg chunk = dS^T Q_h, dVg chunk = P^T dO_h   (16 keys x 32 dims, all 64 queries) ----
        {
          wmma::fragment<wmma::accumulator, 16, 16, 16, float> kacc[2], vacc[2];
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) { wmma::fill_fragment(kacc[nt], 0.f); wmma::fill_fragment(vacc[nt], 0.f); }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> fst, fpt;
            wmma_load_shared(fst, sdSb + ks * 256, 16);
            wmma_load_shared(fpt, sPb + ks * 256, 16);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fq, fg;
              wmma_load_shared(fq, sQ + (ks * 16) * Cfg::LD + hl * 32 + nt * 16, Cfg::LD);
              wmma_load_shared(fg, sdO + (ks * 16) * Cfg::LD + hl * 32 + nt * 16, Cfg::LD);
              wmma::mma_sync(kacc[nt], fst, fq, kacc[nt]);
              wmma::mma_sync(vacc[nt], fpt, fg, vacc[nt]);
            }
          }
          __syncwarp();                                          // every lane is done reading sS / sdP
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            wmma::store_matrix_sync(sS + nt * 16, kacc[nt], 32, wmma::mem_row_major);
            wmma::store_matrix_sync(sdP + nt * 16, vacc[nt], 32, wmma::mem_row_major);
          }
        }
        __syncwarp();
        // ---- bilinear scatter of the 16 keys of this chunk: 8 lanes x 4 channels per key, four keys per pass,
        //      128-bit vector reductions (red.global.add.v4.f32): a quarter of the atomic instructions of a
        //      lane-per-channel scatter, which bound this kernel (1.7 G scalar atomics per launch at config 2) ----
        {
          const int cg = lane & 7, kg = lane >> 3;
          const size_t rbase = (static_cast<size_t>(te) * R + static_cast<size_t>(b * p.L + j) * N) * kC + hgc * 128 + hl * 32 + cg * 4;
          float* dkp = p.dk + rbase;
          float* dvp = p.dv + rbase;
#pragma unroll
          for (int k4 = 0; k4 < 16; k4 += 4) {
            const int kk = k4 + kg;
            const TapRec rec = sTap[kc * 16 + kk];
            if ((rec.w01 | rec.w23) == 0u) continue;
            const float4 gk = *reinterpret_cast<const float4*>(sS + kk * 32 + cg * 4);
            const float4 gv = *reinterpret_cast<const float4*>(sdP + kk * 32 + cg * 4);
            bsum_k4.x += gk.x; bsum_k4.y += gk.y; bsum_k4.z += gk.z; bsum_k4.w += gk.w;
            bsum_v4.x += gv.x; bsum_v4.y += gv.y; bsum_v4.z += gv.z; bsum_v4.w += gv.w;
            const float wq[4] = {bf16_lo(rec.w01), bf16_hi(rec.w01), bf16_lo(rec.w23), bf16_hi(rec.w23)};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (wq[q] == 0.f) continue;
              const size_t off = static_cast<size_t>((rec.y0 + (q >> 1)) * p.W + rec.x0 + (q & 1)) * kC;
#if !(HMVIT_BWD_DBG & 2)
              red_add_v4(dkp + off, wq[q] * gk.x, wq[q] * gk.y, wq[q] * gk.z, wq[q] * gk.w);
              red_add_v4(dvp + off, wq[q] * gv.x, wq[q] * gv.y, wq[q] * gv.z, wq[q] * gv.w);
#endif
            }
          }
        }
        __syncwarp();                                            // scratch is rewritten by the next chunk
      }
      {
        // folded-bias gradients: sum the four key sub-groups (lanes l, l ^ 8, l ^ 16, l ^ 24 hold the same 4 channels)
        float* bk = &bsum_k4.x; float* bv = &bsum_v4.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          bk[e] += __shfl_xor_sync(0xffffffffu, bk[e], 8);  bk[e] += __shfl_xor_sync(0xffffffffu, bk[e], 16);
          bv[e] += __shfl_xor_sync(0xffffffffu, bv[e], 8);  bv[e] += __shfl_xor_sync(0xffffffffu, bv[e], 16);
        }
        if (lane < 8) {
          const int ch4 = hgc * 128 + hl * 32 + lane * 4;
          red_add_v4(p.dbk + (te * 2 + tj) * kC + ch4, bsum_k4.x, bsum_k4.y, bsum_k4.z, bsum_k4.w);
          red_add_v4(p.dbv + (te * 2 + tj) * kC + ch4, bsum_v4.x, bsum_v4.y, bsum_v4.z, bsum_v4.w);
        }
      }
    }
  }

  // ---- dQ: this warp's partial (its key half) -> scratch [64][32] -> fp32 atomics on the ego rows ----
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) wmma::store_matrix_sync(sS + (mt * 16) * 32 + nt * 16, dq[mt][nt], 32, wmma::mem_row_major);
  __syncwarp();
  for (int s4 = 0; s4 < kS; s4 += 4) {
    const int s = s4 + (lane >> 3), cg = lane & 7;
    int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
    const float4 g = *reinterpret_cast<const float4*>(sS + s * 32 + cg * 4);
    red_add_v4(p.dq + (static_cast<size_t>(a) * N + r * p.W + c) * kC + hgc * 128 + hl * 32 + cg * 4, g.x, g.y, g.z, g.w);
  }
  // ---- relative position bias gradient ----
  __syncthreads();
  for (int e = threadIdx.x; e < 225 * kHG; e += Cfg::THREADS) {
    const int idx = e >> 2, h = e & 3;
    const float g = sBgrad[h * kBiasStride + idx] + sBgrad[(h + 4) * kBiasStride + idx];   // the two key-half warps of head h
    if (g != 0.f) atomicAdd(p.dbias_table + idx * kHeads + hgc * kHG + h, g);
  }
}

}  // namespace hmvit
