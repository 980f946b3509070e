// Fused multi-agent group attention for one (scene b, ego agent i, token group g).
//
// Replaces the reference's warp_features + per-ego HeteroAttention.forward core
// (hetero_fusion.py:338-361, 187-277): the warped copies x_pair (B,L,L,C,H,W) and the mask_pair
// tensor are never materialised.  For every source agent j the kernel
//   1. evaluates the j->i source-pixel map for the group's 64 tokens (fp64, like the oracle), the
//      nearest-neighbour ROI visibility (bit-exact key mask) and the 4 bilinear taps;
//   2. gathers the PROJECTED keys / values of agent j (edge-type weights already folded into the
//      projection) with 16-byte vector loads, blends the 4 taps on top of the folded
//      bias with packed bf16 FMAs and stores them swizzled in shared memory;
//   3. runs S = Q K^T (+ relative position bias, key mask), an online softmax over all sources and
//      O += P V on the tensor cores, fp32 accumulation, one warp per 16 query rows x 4 heads.
// Sources with no visible key in this group are skipped entirely (their softmax weight is exactly 0).
// window / grid partition differ only in the token table (hetero_fusion.py:384-389 vs 427-431).
//
// Round-1 note: the contractions use warp-level mma.sync (HMMA) tiles; the kernel is bound by the
// gather (L2 -> SM traffic), see DESIGN.md.  A tcgen05 version is the next step for this kernel.
#pragma once
#include "common.cuh"

namespace hmvit {

#ifdef HMVIT_TS
__device__ unsigned long long g_attn_ts[8][8][4];   // [cta sample][source][event]
#define ATTN_TS(j, ev) do { if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < 8) g_attn_ts[blockIdx.x][j][ev] = clock64(); } while (0)
#else
#define ATTN_TS(j, ev) do { } while (0)
#endif

struct AttnParams {
  int B, L, H, W;
  int kind;                    // 0 = window partition, 1 = grid partition
  int ego_only;                // 1: only ego slot 0 of each scene
  const int* mode;             // [B*L]
  const int* record_len;       // [B]
  const int* cav_mask;         // [B*L]
  const float* T;              // [B][L][L][16]  pairwise_t_matrix, [b][j][i] maps j -> i
  double cell;                 // voxel_size[0] * downsample_rate (metres per BEV cell)
  const __nv_bfloat16* q;      // [B*L*N][256]
  const __nv_bfloat16* k;      // [2 (te)][B*L*N][256]
  const __nv_bfloat16* v;      // [2 (te)][B*L*N][256]
  const float* bk;             // [2 (te)][2 (tj)][256] folded key bias (added after the gather)
  const float* bv;             // [2 (te)][2 (tj)][256]
  const float* bias_table;     // [225][8] relative_position_bias_table.weight
  const uint8_t* key_mask;     // optional [B*L][N]: extra per-token key mask of source j (unit-level API), or null
  __nv_bfloat16* out;          // [B*L*N][256]
  float* lse;                  // optional [B*L*N][8]: log2-domain log-sum-exp of every (query row, head), saved for the backward
};

constexpr int kAttnThreads = 128;
constexpr int kHG = 4;                           // heads per CTA (a CTA owns one head group = 128 channels of one group of tokens)
constexpr int kRowBytes = kHG * kDh * 2;         // 256 B: one token row of the head group
constexpr int kTileBytes = kS * kRowBytes;       // 16 KB: 64 tokens x 128 ch bf16
constexpr int kBiasStride = 232;                 // [4 heads][225 (+7 pad)]
// per (source, token) gather record, 12 bytes: corner (x0, y0) and the 4 tap weights as bf16.  A token is
// visible iff some weight is non-zero (the nearest in-range corner always has weight >= 1/4).
struct TapRec { short x0, y0; uint32_t w01; uint32_t w23; };     // w01 = (w00, w01), w23 = (w10, w11)
constexpr int kMaxSrc = 5;                       // sources per tap pass
constexpr int kAttnSmem = 3 * kTileBytes + kHG * kBiasStride * 4 + kMaxSrc * kS * sizeof(TapRec) + 32;

// element (row, 16-byte unit 0..15) of a [64][256 B] tile, XOR-swizzled so ldmatrix is conflict free
HMVIT_DEVINL uint32_t tile_off(int row, int unit) { return row * kRowBytes + ((unit ^ (row & 7)) << 4); }

HMVIT_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HMVIT_DEVINL void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
HMVIT_DEVINL void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
HMVIT_DEVINL float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
HMVIT_DEVINL uint32_t hfma2_bf16(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// flat token index of slot s in group (gy, gx)
HMVIT_DEVINL void group_token(int kind, int gy, int gx, int s, int H, int W, int& r, int& c) {
  const int s1 = s >> 3, s2 = s & 7;
  if (kind == 0) { r = gy * kWin + s1; c = gx * kWin + s2; }
  else           { r = s1 * (H / kWin) + gy; c = s2 * (W / kWin) + gx; }
}

// One CTA = (scene b, ego i, group of 64 tokens, head group of 4 heads); 4 warps, warp w owns query rows
// [16w, 16w+16) x 4 heads.  Four CTAs are resident per SM (55 KB of shared memory, 128 registers), so
// gather phases (L2 latency) of some overlap the tensor-core / softmax phases of the others.
__global__ void __launch_bounds__(kAttnThreads, 4) group_attn_kernel(const AttnParams p) {
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);            // a malformed record_len must not index past the scene's slots
  if (i >= nrec || (p.ego_only && i != 0)) return;
  const int N = p.H * p.W;
  const int GX = p.W / kWin;
  const int grp = blockIdx.x >> 1, hgc = blockIdx.x & 1;      // token group, head group
  const int gy = grp / GX, gx = grp - gy * GX;
  const int te = p.mode[a] != 0 ? 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = warp;
  const int g = lane >> 2, t = lane & 3;
  const int hl = lane >> 4, u16 = lane & 15;                  // half-warp (one token each) and 16-byte unit within the 256-byte row

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 2 * kTileBytes;
  float* sBias = reinterpret_cast<float*>(smem + 3 * kTileBytes);             // [4][kBiasStride], log2 domain
  TapRec* sTapAll = reinterpret_cast<TapRec*>(sBias + kHG * kBiasStride);     // [kMaxSrc][64]
  int* sAnyVis = reinterpret_cast<int*>(sTapAll + kMaxSrc * kS);              // [kMaxSrc]

  const int cu0 = hgc * 16;                                    // first 16-byte unit of this head group in a 512-byte row
  // ---- stage Q (ego rows; softmax scale and log2(e) are folded into W_q) and the bias table ----
  {
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
#pragma unroll 4
    for (int tt = 0; tt < 8; ++tt) {
      const int s = warp * 16 + tt * 2 + hl;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      const uint4 v = __ldg(qsrc + static_cast<size_t>(r * p.W + c) * 32);
      *reinterpret_cast<uint4*>(sQ + tile_off(s, u16)) = v;
    }
    for (int e = threadIdx.x; e < 225 * kHG; e += kAttnThreads) {
      const int idx = e >> 2, h = e & 3;
      sBias[h * kBiasStride + idx] = __ldg(p.bias_table + idx * kHeads + hgc * kHG + h) * 1.4426950408889634f;
    }
  }

  float o[kHG][4][4];
  float mrow[kHG][2], lrow[kHG][2];
#pragma unroll
  for (int h = 0; h < kHG; ++h) {
    mrow[h][0] = mrow[h][1] = -INFINITY; lrow[h][0] = lrow[h][1] = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[h][n][0] = o[h][n][1] = o[h][n][2] = o[h][n][3] = 0.f; }
  }

  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
  const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;            // uint4 units per te plane
  // relative-position bias index of (query row g | g+8 of block rb, key nt*8 + 2t + e):
  //   ((2rb [+1]) - nt + 7) * 15 + (g - (2t + e) + 7)  =  bias_base - 15 nt - e  [+ 15]
  const int bias_base = (2 * rb + 7) * 15 + (g - 2 * t + 7);

  for (int j0 = 0; j0 < nrec; j0 += kMaxSrc) {
  const int nsrc = min(kMaxSrc, nrec - j0);
  // ---- taps + visibility of every (source, token) of this group, all sources in one parallel pass ----
  __syncthreads();                               // previous pass (and the Q / bias staging) done with smem
  ATTN_TS(j0, 0);
  if (threadIdx.x < kMaxSrc) sAnyVis[threadIdx.x] = 0;
  __syncthreads();
  for (int e = threadIdx.x; e < nsrc * kS; e += kAttnThreads) {
    const int js = e >> 6, tk = e & 63, j = j0 + js;
    TapRec rec; rec.x0 = 0; rec.y0 = 0; rec.w01 = 0; rec.w23 = 0;
    if (p.cav_mask[b * p.L + j] != 0) {
      const WarpMap wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + j) * p.L + i) * 16, p.H, p.W, p.cell);
      int r, c; group_token(p.kind, gy, gx, tk, p.H, p.W, r, c);
      double sx, sy; warp_src(wm, c, r, sx, sy);
      bool vis = warp_visible(sx, sy, p.H, p.W);
      if (p.key_mask != nullptr && p.key_mask[static_cast<size_t>(b * p.L + j) * N + r * p.W + c] == 0) vis = false;
      if (vis) {
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
    if (sAnyVis[js] == 0) continue;              // source invisible in this group: its softmax weight is exactly 0
    const TapRec* sTap = sTapAll + js * kS;
    ATTN_TS(j, 1);

    // ---- gather projected K / V rows of source j (this head group's 256 bytes): 4-tap blend with packed
    //      bf16 FMAs on top of the folded bias; a half-warp per token ----
    {
      const int tj = p.mode[b * p.L + j] != 0 ? 1 : 0;
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
#pragma unroll 2
      for (int tt = 0; tt < 8; ++tt) {
        const int s = warp * 16 + tt * 2 + hl;
        const TapRec rec = sTap[s];
        uint4 ko = make_uint4(0, 0, 0, 0), vo = make_uint4(0, 0, 0, 0);
        if ((rec.w01 | rec.w23) != 0u) {
          // bf16 weights; a tap that falls outside the map has weight 0 and is not loaded
          const uint32_t wq[4] = {rec.w01 & 0xffffu, rec.w01 >> 16, rec.w23 & 0xffffu, rec.w23 >> 16};
          uint4 kk[4], vv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int yy = rec.y0 + (q >> 1), xx = rec.x0 + (q & 1);
            kk[q] = make_uint4(0, 0, 0, 0); vv[q] = make_uint4(0, 0, 0, 0);
            if (wq[q] != 0u) {
              const size_t off = static_cast<size_t>(yy * p.W + xx) * 32;
              kk[q] = __ldg(ksrc + off); vv[q] = __ldg(vsrc + off);
            }
          }
          ko = make_uint4(bk2[0], bk2[1], bk2[2], bk2[3]);
          vo = make_uint4(bv2[0], bv2[1], bv2[2], bv2[3]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t w2 = wq[q] | (wq[q] << 16);
            ko.x = hfma2_bf16(w2, kk[q].x, ko.x); ko.y = hfma2_bf16(w2, kk[q].y, ko.y);
            ko.z = hfma2_bf16(w2, kk[q].z, ko.z); ko.w = hfma2_bf16(w2, kk[q].w, ko.w);
            vo.x = hfma2_bf16(w2, vv[q].x, vo.x); vo.y = hfma2_bf16(w2, vv[q].y, vo.y);
            vo.z = hfma2_bf16(w2, vv[q].z, vo.z); vo.w = hfma2_bf16(w2, vv[q].w, vo.w);
          }
        }
        *reinterpret_cast<uint4*>(sK + tile_off(s, u16)) = ko;
        *reinterpret_cast<uint4*>(sV + tile_off(s, u16)) = vo;
      }
    }
    __syncthreads();
    ATTN_TS(j, 2);

    // ---- key visibility bits for this thread's columns: key s' = nt*8 + 2t + e -> bit nt*2 + e ----
    uint32_t vbits = 0;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const TapRec r0 = sTap[nt * 8 + 2 * t], r1 = sTap[nt * 8 + 2 * t + 1];
      vbits |= ((r0.w01 | r0.w23) != 0u ? 1u : 0u) << (nt * 2);
      vbits |= ((r1.w01 | r1.w23) != 0u ? 1u : 0u) << (nt * 2 + 1);
    }

    // ---- tensor-core phase: this warp's 16 query rows x 4 heads ----
#pragma unroll
    for (int hh = 0; hh < kHG; ++hh) {
      // accumulators start from the relative position bias (log2 domain)
      const float* bh = sBias + hh * kBiasStride + bias_base;
      float sacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sacc[nt][0] = bh[-15 * nt];      sacc[nt][1] = bh[-15 * nt - 1];
        sacc[nt][2] = bh[-15 * nt + 15]; sacc[nt][3] = bh[-15 * nt + 14];
      }
#pragma unroll
      for (int kk2 = 0; kk2 < 2; ++kk2) {                       // 2 x k16 over the 32 head dims
        uint32_t a0, a1, a2, a3;
        {
          const int row = rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int unit = hh * 4 + kk2 * 2 + (lane >> 4);
          ldsm_x4(sQ_u + tile_off(row, unit), a0, a1, a2, a3);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {                        // pairs of key n-tiles
          uint32_t b0, b1, b2, b3;
          const int row = np * 16 + (lane & 7) + (lane >> 4) * 8;
          const int unit = hh * 4 + kk2 * 2 + ((lane >> 3) & 1);
          ldsm_x4(sK_u + tile_off(row, unit), b0, b1, b2, b3);
          mma_bf16(sacc[np * 2], a0, a1, a2, a3, b0, b1);
          mma_bf16(sacc[np * 2 + 1], a0, a1, a2, a3, b2, b3);
        }
      }
      // key mask, running max
      float mx0 = -INFINITY, mx1 = -INFINITY;
      if (vbits != 0xffffu) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (!((vbits >> (nt * 2 + e)) & 1u)) { sacc[nt][e] = -INFINITY; sacc[nt][2 + e] = -INFINITY; }
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(mrow[hh][0], mx0), mn1 = fmaxf(mrow[hh][1], mx1);
      const float mu0 = (mn0 == -INFINITY) ? 0.f : mn0, mu1 = (mn1 == -INFINITY) ? 0.f : mn1;
      const float al0 = ex2(mrow[hh][0] - mu0), al1 = ex2(mrow[hh][1] - mu1);
      mrow[hh][0] = mn0; mrow[hh][1] = mn1;
      uint32_t pa[4][4];                                         // P as A fragments: 4 x k16 over the 64 keys
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = ex2(sacc[nt][0] - mu0), p1 = ex2(sacc[nt][1] - mu0);
        const float p2 = ex2(sacc[nt][2] - mu1), p3 = ex2(sacc[nt][3] - mu1);
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
      }
      // running denominator: l = l * alpha + rowsum(P), the row sums come from the tensor core too
      // (P x ones), i.e. from exactly the bf16 probabilities that multiply V
      float lacc[4] = {lrow[hh][0] * al0, 0.f, lrow[hh][1] * al1, 0.f};
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) mma_bf16(lacc, pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], 0x3F803F80u, 0x3F803F80u);
      lrow[hh][0] = lacc[0]; lrow[hh][1] = lacc[2];
#pragma unroll
      for (int n = 0; n < 4; ++n) { o[hh][n][0] *= al0; o[hh][n][1] *= al0; o[hh][n][2] *= al1; o[hh][n][3] *= al1; }
      // O_h += P V_h
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
#pragma unroll
        for (int np = 0; np < 2; ++np) {                        // pairs of dim n-tiles
          uint32_t b0, b1, b2, b3;
          const int row = kc * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          const int unit = hh * 4 + np * 2 + (lane >> 4);
          ldsm_x4_t(sV_u + tile_off(row, unit), b0, b1, b2, b3);
          mma_bf16(o[hh][np * 2], pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], b0, b1);
          mma_bf16(o[hh][np * 2 + 1], pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], b2, b3);
        }
      }
    }
    __syncthreads();     // compute phase done before sK / sV are rewritten for the next source
    ATTN_TS(j, 3);
  }
  }

  // ---- training: save the softmax statistics (log2 domain: running max + log2 of the denominator) ----
  if (p.lse != nullptr && t == 0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      int r, c; group_token(p.kind, gy, gx, rb * 16 + g + e * 8, p.H, p.W, r, c);
      float* dst = p.lse + (static_cast<size_t>(a) * N + r * p.W + c) * kHeads + hgc * kHG;
#pragma unroll
      for (int hh = 0; hh < kHG; ++hh) dst[hh] = lrow[hh][e] > 0.f ? mrow[hh][e] + log2f(lrow[hh][e]) : INFINITY;
    }
  }
  // ---- normalise, stage in smem (reuse sK), coalesced store ----
  __syncthreads();
#pragma unroll
  for (int hh = 0; hh < kHG; ++hh) {
    const float il0 = lrow[hh][0] > 0.f ? 1.0f / lrow[hh][0] : 0.f;
    const float il1 = lrow[hh][1] > 0.f ? 1.0f / lrow[hh][1] : 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col = hh * kDh + n * 8 + 2 * t;                  // channel within the head group
      const int r0 = rb * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(sK + tile_off(r0, col >> 3) + (col & 7) * 2) = pack_bf16x2(o[hh][n][0] * il0, o[hh][n][1] * il0);
      *reinterpret_cast<uint32_t*>(sK + tile_off(r1, col >> 3) + (col & 7) * 2) = pack_bf16x2(o[hh][n][2] * il1, o[hh][n][3] * il1);
    }
  }
  __syncthreads();
  {
    uint4* dst = reinterpret_cast<uint4*>(p.out) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
#pragma unroll 4
    for (int tt = 0; tt < 8; ++tt) {
      const int s = warp * 16 + tt * 2 + hl;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      dst[static_cast<size_t>(r * p.W + c) * 32] = *reinterpret_cast<const uint4*>(sK + tile_off(s, u16));
    }
  }
}

}  // namespace hmvit
