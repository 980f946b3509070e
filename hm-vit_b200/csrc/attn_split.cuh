// Split form of the fused warp + mask + multi-agent group attention (same contract as attn.cuh):
//
//   warp_compact_kernel   one CTA per (scene b, ego i, token group g).  Evaluates the j -> i source-pixel map
//                         of every (source j, token) of the group (fp64, bit-exact ROI visibility like
//                         attn.cuh), COMPACTS the visible (source, token) pairs of all sources into one key
//                         list, gathers + blends their projected K' / V' rows (4 taps, packed bf16 FMAs on top
//                         of the folded bias) and writes them as dense 64-key tiles, already in the swizzled
//                         shared-memory image the attention kernel wants (one 16 KB blob per tile x head group).
//                         Small CTAs, few registers -> many warps per SM, which is what the tap gather needs
//                         (profiles/r1_attention_study.md: gather bandwidth scales with gathering warps per SM).
//   dense_attn_kernel     one CTA per (b, i, g, head group of 4 heads), warp == head.  Per 64-key tile: two 16 KB
//                         bulk copies (cp.async.bulk -> mbarrier, K and V double buffered), the tile's K / V
//                         fragments read from shared memory ONCE and held in registers for the four 16-row query
//                         blocks, S = Q K^T + relative position bias (looked up through the tile's key-slot list,
//                         prefetched one tile ahead), online softmax, O += P V with mma.sync bf16 tiles.
//                         No key mask inside a tile: only the tail of the last tile is masked.
//
// Replaces hetero_fusion.py:338-361 (warp_features) + :187-277 (HeteroAttention.forward core).  Compared with
// the single-kernel form the key tiles make a round trip through HBM (visible keys only, bf16), but the
// grid partition needs 37 % fewer key tiles (keys of different sources share a tile) and neither half waits
// on the other's latency.
#pragma once
#include "attn.cuh"

namespace hmvit {

constexpr int kSplitMaxL = 8;                              // agents per scene the compaction pass handles
constexpr int kBlobBytes = kTileBytes;                     // 16 KB: 64 keys x 128 channels bf16 (one head group)
constexpr int kCompactThreads = 256;

struct SplitParams {
  AttnParams a;
  uint8_t* kc;                 // [B*L][G][L][2 hg][16 KB] compacted, blended keys (swizzled tile image)
  uint8_t* vc;                 // same for values
  int* nvis;                   // [B*L][G] visible keys of the (ego, group)
  uint8_t* slots;              // [B*L][G][L*64] group slot (0..63) of every compacted key
  int tc_layout;               // tile image: 0 = [64 keys][256 B] XOR-swizzled (ldmatrix, dense_attn_kernel);
                               // 1 = [2 head pairs][64 keys][128 B] UMMA SWIZZLE_128B (dense_attn_tc_kernel)
};

struct CRec { short x0, y0; uint32_t w01, w23, meta; };    // meta = source j | slot << 8

// The ego's own keys need no warp when T[b][i][i] is exactly the identity (pairwise_t_matrix is built that way,
// intermediate_fusion_dataset.py:181-200): the dense pass then reads them straight from the projected rows and the
// compaction pass leaves them out.  Any other diagonal goes through the general path like every other source.
HMVIT_DEVINL bool self_is_identity(const AttnParams& p, int b, int i) {
  if (p.key_mask != nullptr) return false;
  const float* t = p.T + ((static_cast<size_t>(b) * p.L + i) * p.L + i) * 16;
  return t[0] == 1.f && t[1] == 0.f && t[3] == 0.f && t[4] == 0.f && t[5] == 1.f && t[7] == 0.f && p.cav_mask[b * p.L + i] != 0;
}

__global__ void __launch_bounds__(kCompactThreads, 4) warp_compact_kernel(const SplitParams sp) {
  const AttnParams& p = sp.a;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);            // a malformed record_len must not index past the scene's slots
  if (i >= nrec || (p.ego_only && i != 0)) return;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int grp = blockIdx.x;
  const int gy = grp / GX, gx = grp - gy * GX;
  const int te = p.mode[a] != 0 ? 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  __shared__ __align__(16) CRec sRec[kSplitMaxL * kS];
  __shared__ int sCnt[kSplitMaxL * 2];
  __shared__ int sTj[kSplitMaxL];

  if (threadIdx.x < kSplitMaxL) sTj[threadIdx.x] = (threadIdx.x < nrec && p.mode[b * p.L + threadIdx.x] != 0) ? 1 : 0;
  const int jskip = self_is_identity(p, b, i) ? i : -1;

  // ---- pass 1: taps + visibility of every (source, token), identical arithmetic to attn.cuh ----
  const int nent = nrec * kS;
  CRec rec[2];
  uint32_t bal[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int e = it * kCompactThreads + threadIdx.x;
    CRec r; r.x0 = 0; r.y0 = 0; r.w01 = 0; r.w23 = 0; r.meta = 0;
    bool vis = false;
    if (e < nent) {
      const int j = e >> 6, tk = e & 63;
      r.meta = static_cast<uint32_t>(j) | (static_cast<uint32_t>(tk) << 8);
      if (j != jskip && p.cav_mask[b * p.L + j] != 0) {
        const WarpMap wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + j) * p.L + i) * 16, p.H, p.W, p.cell);
        int rr, cc; group_token(p.kind, gy, gx, tk, p.H, p.W, rr, cc);
        double sx, sy; warp_src(wm, cc, rr, sx, sy);
        vis = warp_visible(sx, sy, p.H, p.W);
        if (p.key_mask != nullptr && p.key_mask[static_cast<size_t>(b * p.L + j) * N + rr * p.W + cc] == 0) vis = false;
        if (vis) {
          const Taps tp = make_taps(sx, sy, p.H, p.W);
          r.x0 = static_cast<short>(tp.x0); r.y0 = static_cast<short>(tp.y0);
          r.w01 = pack_bf16x2(tp.w00, tp.w01); r.w23 = pack_bf16x2(tp.w10, tp.w11);
          // a visible key always has a non-zero tap weight (the nearest in-range corner has weight >= 1/4)
        }
      }
    }
    rec[it] = r;
    bal[it] = __ballot_sync(0xffffffffu, vis);
    if (lane == 0) sCnt[it * 8 + warp] = __popc(bal[it]);
  }
  __syncthreads();
  int nv = 0;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int seg = it * 8 + warp;
    int base = 0;
#pragma unroll
    for (int s = 0; s < 2 * kSplitMaxL; ++s) {
      const int c = sCnt[s];
      if (s < seg) base += c;
      nv += (it == 0) ? c : 0;
    }
    if ((bal[it] >> lane) & 1u) sRec[base + __popc(bal[it] & ((1u << lane) - 1u))] = rec[it];
  }
  __syncthreads();
  const int nfill = (nv + kS - 1) & ~(kS - 1);            // rows written: visible keys + zero tail of the last tile

  // ---- meta data for the attention kernel ----
  const size_t ag = static_cast<size_t>(a) * G + grp;
  if (threadIdx.x == 0) sp.nvis[ag] = nv;
  {
    uint8_t* sl = sp.slots + ag * (p.L * kS);
    for (int e = threadIdx.x; e < nfill; e += kCompactThreads) sl[e] = e < nv ? static_cast<uint8_t>(sRec[e].meta >> 8) : 0;
  }

  // ---- pass 2: gather + blend; a warp per key row (512 B = both head groups), 8 loads in flight per lane ----
  const int hg = lane >> 4, u16 = lane & 15;
  uint32_t bkq[2][4], bvq[2][4];
#pragma unroll
  for (int tj = 0; tj < 2; ++tj) {
    const float4* pk = reinterpret_cast<const float4*>(p.bk + (te * 2 + tj) * kC + lane * 8);
    const float4* pv = reinterpret_cast<const float4*>(p.bv + (te * 2 + tj) * kC + lane * 8);
    const float4 k0 = __ldg(pk), k1 = __ldg(pk + 1), v0 = __ldg(pv), v1 = __ldg(pv + 1);
    bkq[tj][0] = pack_bf16x2(k0.x, k0.y); bkq[tj][1] = pack_bf16x2(k0.z, k0.w); bkq[tj][2] = pack_bf16x2(k1.x, k1.y); bkq[tj][3] = pack_bf16x2(k1.z, k1.w);
    bvq[tj][0] = pack_bf16x2(v0.x, v0.y); bvq[tj][1] = pack_bf16x2(v0.z, v0.w); bvq[tj][2] = pack_bf16x2(v1.x, v1.y); bvq[tj][3] = pack_bf16x2(v1.z, v1.w);
  }
  const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;            // uint4 units per te plane
  const uint4* kbase = reinterpret_cast<const uint4*>(p.k) + te * plane + static_cast<size_t>(b) * p.L * N * 32 + lane;
  const uint4* vbase = reinterpret_cast<const uint4*>(p.v) + te * plane + static_cast<size_t>(b) * p.L * N * 32 + lane;
  uint8_t* kdst = sp.kc + ag * (static_cast<size_t>(p.L) * 2 * kBlobBytes);
  uint8_t* vdst = sp.vc + ag * (static_cast<size_t>(p.L) * 2 * kBlobBytes);
  constexpr int kWarps = kCompactThreads / 32;
  for (int pos = warp; pos < nfill; pos += kWarps) {
    uint4 ko = make_uint4(0, 0, 0, 0), vo = make_uint4(0, 0, 0, 0);
    if (pos < nv) {
      const CRec r = sRec[pos];
      const int j = r.meta & 0xff;
      const int tj = sTj[j];
      const uint32_t wq[4] = {r.w01 & 0xffffu, r.w01 >> 16, r.w23 & 0xffffu, r.w23 >> 16};
      uint4 kk[4], vv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        kk[q] = make_uint4(0, 0, 0, 0); vv[q] = make_uint4(0, 0, 0, 0);
        if (wq[q] != 0u) {          // a tap that falls outside the map has weight 0 and is not loaded
          const size_t off = (static_cast<size_t>(j) * N + (r.y0 + (q >> 1)) * p.W + r.x0 + (q & 1)) * 32;
          kk[q] = __ldg(kbase + off); vv[q] = __ldg(vbase + off);
        }
      }
      ko = tj ? make_uint4(bkq[1][0], bkq[1][1], bkq[1][2], bkq[1][3]) : make_uint4(bkq[0][0], bkq[0][1], bkq[0][2], bkq[0][3]);
      vo = tj ? make_uint4(bvq[1][0], bvq[1][1], bvq[1][2], bvq[1][3]) : make_uint4(bvq[0][0], bvq[0][1], bvq[0][2], bvq[0][3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t w2 = wq[q] | (wq[q] << 16);
        ko.x = hfma2_bf16(w2, kk[q].x, ko.x); ko.y = hfma2_bf16(w2, kk[q].y, ko.y);
        ko.z = hfma2_bf16(w2, kk[q].z, ko.z); ko.w = hfma2_bf16(w2, kk[q].w, ko.w);
        vo.x = hfma2_bf16(w2, vv[q].x, vo.x); vo.y = hfma2_bf16(w2, vv[q].y, vo.y);
        vo.z = hfma2_bf16(w2, vv[q].z, vo.z); vo.w = hfma2_bf16(w2, vv[q].w, vo.w);
      }
    }
    const size_t o = static_cast<size_t>((pos >> 6) * 2 + hg) * kBlobBytes +
                     (sp.tc_layout ? (u16 >> 3) * 8192 + sw128_offset(pos & 63, u16 & 7) : tile_off(pos & 63, u16));
    *reinterpret_cast<uint4*>(kdst + o) = ko;
    *reinterpret_cast<uint4*>(vdst + o) = vo;
  }
}

// ------------------------------------------------------------------------------------------------
// shared memory: Q | K buffer 0 | K buffer 1 | V buffer 0 | V buffer 1 | bias table | 4 mbarriers  (84 KB)
constexpr int kDenseSmem = 5 * kTileBytes + kHG * kBiasStride * 4 + 32;
#ifndef HMVIT_DENSE_WARPS
#define HMVIT_DENSE_WARPS 4                                // 4 query row blocks x (warps / 4) head sets; measured on B200
                                                           // (config 2): 4 warps x 3 CTAs/SM 0.48-0.53 ms, 8 x 2 0.50-0.55 ms
#endif
#ifndef HMVIT_DENSE_HEADWARP
#define HMVIT_DENSE_HEADWARP 1                             // 1: warp == head (all 64 query rows), K / V fragments read from shared
                                                           // memory once per tile instead of once per 16-row block (4 warps only):
                                                           // 46 % fewer shared-memory wavefronts, 255 registers, 2 CTAs / SM;
                                                           // measured 0.353 vs 0.365 ms (0: warp == 16-row block x 4 heads, 3 CTAs / SM)
#endif
#ifndef HMVIT_DENSE_CTAS
#define HMVIT_DENSE_CTAS (HMVIT_DENSE_HEADWARP ? 2 : 3)                                 // resident CTAs per SM the register budget is sized for
#endif
constexpr int kDenseWarps = HMVIT_DENSE_WARPS;
constexpr int kDenseThreads = kDenseWarps * 32;
constexpr int kHPW = kHG / (kDenseWarps / 4);              // heads per warp
constexpr int kStageIt = 8 / (kDenseWarps / 4);            // token pairs a warp stages (Q, self tile, output)
constexpr int kBiasIt = (225 * kHG + kDenseThreads - 1) / kDenseThreads;
static_assert(kDenseWarps == 4 || kDenseWarps == 8 || kDenseWarps == 16, "dense attention: 4, 8 or 16 warps");
static_assert(!HMVIT_DENSE_HEADWARP || kDenseWarps == 4, "warp == head needs 4 warps");

HMVIT_DEVINL void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Key tiles are consumed in order [self tile (when the ego's own transform is the identity)] [compacted tiles 0..].
// K and V are double buffered: the next tile's keys and values arrive during this tile's work.
__global__ void __launch_bounds__(kDenseThreads, HMVIT_DENSE_CTAS) dense_attn_kernel(const SplitParams sp) {
  const AttnParams& p = sp.a;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);            // a malformed record_len must not index past the scene's slots
  if (i >= nrec || (p.ego_only && i != 0)) return;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int grp = blockIdx.x >> 1, hgc = blockIdx.x & 1;      // token group, head group
  const int gy = grp / GX, gx = grp - gy * GX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = warp & 3, h0 = (warp >> 2) * kHPW;            // query row block, first head (of the head group) of this warp
  const int g = lane >> 2, t = lane & 3;
  const int hl = lane >> 4, u16 = lane & 15;

  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;                                            // two buffers
  uint8_t* sV = smem + 3 * kTileBytes;
  float* sBias = reinterpret_cast<float*>(smem + 5 * kTileBytes);             // [4][kBiasStride], log2 domain
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + kHG * kBiasStride);    // K buffer 0, K buffer 1, V buffer 0, V buffer 1

  const size_t ag = static_cast<size_t>(a) * G + grp;
  const int nv = sp.nvis[ag];
  const int ntiles = (nv + kS - 1) >> 6;
  const int self = self_is_identity(p, b, i) ? 1 : 0;
  const uint8_t* kblob = sp.kc + ag * (static_cast<size_t>(p.L) * 2 * kBlobBytes) + hgc * kBlobBytes;
  const uint8_t* vblob = sp.vc + ag * (static_cast<size_t>(p.L) * 2 * kBlobBytes) + hgc * kBlobBytes;
  const uint8_t* slots = sp.slots + ag * (p.L * kS);
  const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);

  if (threadIdx.x == 0) {
    mbar_init(bars + 0, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1); mbar_init(bars + 3, 1);
    fence_mbar_init();
    if (ntiles > 0) {                                          // compacted tile 0 -> buffers `self` (the self tile, if any, owns buffers 0)
      mbar_arrive_expect_tx(bars + self, kBlobBytes);
      bulk_load(sK_u + self * kTileBytes, kblob, kBlobBytes, bars + self);
      mbar_arrive_expect_tx(bars + 2 + self, kBlobBytes);
      bulk_load(sV_u + self * kTileBytes, vblob, kBlobBytes, bars + 2 + self);
    }
  }

  const int cu0 = hgc * 16;
  // ---- stage Q (ego rows; softmax scale and log2(e) are folded into W_q), the bias table, and the self tile.
  //      All global loads are issued before the first shared-memory store: with 12 warps per SM nothing else
  //      hides this latency ----
  {
    const int te = p.mode[a] != 0 ? 1 : 0;
    const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;
    const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    const uint4* ksrc = reinterpret_cast<const uint4*>(p.k) + te * plane + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    const uint4* vsrc = reinterpret_cast<const uint4*>(p.v) + te * plane + static_cast<size_t>(a) * N * 32 + cu0 + u16;
    uint4 qv[kStageIt], kv[kStageIt], vv[kStageIt];
    float bt[kBiasIt];
#pragma unroll
    for (int tt = 0; tt < kStageIt; ++tt) {
      const int s = warp * (2 * kStageIt) + tt * 2 + hl;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      const size_t off = static_cast<size_t>(r * p.W + c) * 32;
      qv[tt] = __ldg(qsrc + off);
      if (self) { kv[tt] = __ldg(ksrc + off); vv[tt] = __ldg(vsrc + off); }
    }
#pragma unroll
    for (int it = 0; it < kBiasIt; ++it) {
      const int e = it * kDenseThreads + threadIdx.x;
      bt[it] = e < 225 * kHG ? __ldg(p.bias_table + (e >> 2) * kHeads + hgc * kHG + (e & 3)) : 0.f;
    }
    uint32_t bk2[4] = {0, 0, 0, 0}, bv2[4] = {0, 0, 0, 0};
    if (self) {
      // own keys / values: one tap of weight 1 on top of the folded bias -- the arithmetic of the general gather
      const float4* pk = reinterpret_cast<const float4*>(p.bk + (te * 2 + te) * kC + (cu0 + u16) * 8);
      const float4* pv = reinterpret_cast<const float4*>(p.bv + (te * 2 + te) * kC + (cu0 + u16) * 8);
      const float4 k0 = __ldg(pk), k1 = __ldg(pk + 1), v0 = __ldg(pv), v1 = __ldg(pv + 1);
      bk2[0] = pack_bf16x2(k0.x, k0.y); bk2[1] = pack_bf16x2(k0.z, k0.w); bk2[2] = pack_bf16x2(k1.x, k1.y); bk2[3] = pack_bf16x2(k1.z, k1.w);
      bv2[0] = pack_bf16x2(v0.x, v0.y); bv2[1] = pack_bf16x2(v0.z, v0.w); bv2[2] = pack_bf16x2(v1.x, v1.y); bv2[3] = pack_bf16x2(v1.z, v1.w);
    }
    constexpr uint32_t kOne2 = 0x3F803F80u;
#pragma unroll
    for (int tt = 0; tt < kStageIt; ++tt) {
      const int s = warp * (2 * kStageIt) + tt * 2 + hl;
      *reinterpret_cast<uint4*>(sQ + tile_off(s, u16)) = qv[tt];
      if (self) {
        uint4 ko, vo;
        ko.x = hfma2_bf16(kOne2, kv[tt].x, bk2[0]); ko.y = hfma2_bf16(kOne2, kv[tt].y, bk2[1]);
        ko.z = hfma2_bf16(kOne2, kv[tt].z, bk2[2]); ko.w = hfma2_bf16(kOne2, kv[tt].w, bk2[3]);
        vo.x = hfma2_bf16(kOne2, vv[tt].x, bv2[0]); vo.y = hfma2_bf16(kOne2, vv[tt].y, bv2[1]);
        vo.z = hfma2_bf16(kOne2, vv[tt].z, bv2[2]); vo.w = hfma2_bf16(kOne2, vv[tt].w, bv2[3]);
        *reinterpret_cast<uint4*>(sK + tile_off(s, u16)) = ko;
        *reinterpret_cast<uint4*>(sV + tile_off(s, u16)) = vo;
      }
    }
#pragma unroll
    for (int it = 0; it < kBiasIt; ++it) {
      const int e = it * kDenseThreads + threadIdx.x;
      if (e < 225 * kHG) sBias[(e & 3) * kBiasStride + (e >> 2)] = bt[it] * 1.4426950408889634f;
    }
  }

  float o[kHPW][4][4];
  float mrow[kHPW][2], lrow[kHPW][2];
#pragma unroll
  for (int h = 0; h < kHPW; ++h) {
    mrow[h][0] = mrow[h][1] = -INFINITY; lrow[h][0] = lrow[h][1] = 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) { o[h][n][0] = o[h][n][1] = o[h][n][2] = o[h][n][3] = 0.f; }
  }
  // relative-position bias of (query row g | g+8 of block rb, key slot s'):
  //   ((2rb [+1]) - (s' >> 3) + 7) * 15 + (g - (s' & 7) + 7)  =  bias_q - koff(s')  [+ 15],  koff = (s' >> 3) * 15 + (s' & 7)
  const int bias_q = (2 * rb + 7) * 15 + (g + 7);
  __syncthreads();                                             // Q / bias / self tile staged, barriers initialised

  // key slots of this thread's 16 columns of a compacted tile (keys nt*8 + 2t + e), requested one tile ahead
  uint32_t slot_pf[4] = {0, 0, 0, 0};
  auto load_slots = [&](int tile_) {
    const uint16_t* sl = reinterpret_cast<const uint16_t*>(slots + tile_ * kS) + t;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      slot_pf[q] = __ldg(sl + (2 * q) * 4) | (static_cast<uint32_t>(__ldg(sl + (2 * q + 1) * 4)) << 16);   // key pairs nt = 2q, 2q+1
  };
  if (ntiles > 0) load_slots(0);
  uint32_t phases = 0;                                         // bit n: parity the next wait on barrier n expects
  const int nvt = self + ntiles;
  for (int vt = 0; vt < nvt; ++vt) {
    const int tile = vt - self;                                // -1: the self tile
    const int kb = vt & 1;
    // key slots of this thread's 16 columns (key nt*8 + 2t + e): bias offsets packed 4 x 8 bit
    uint32_t koff[4];
    if (tile < 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t k00 = (2 * q) * 15 + 2 * t, k10 = (2 * q + 1) * 15 + 2 * t;
        koff[q] = k00 | ((k00 + 1) << 8) | (k10 << 16) | ((k10 + 1) << 24);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t s4 = slot_pf[q];                                                    // 4 slot bytes (prefetched one tile ahead)
        koff[q] = ((s4 >> 3) & 0x07070707u) * 15u + (s4 & 0x07070707u);                    // per byte, <= 112: no carry
      }
      if (tile + 1 < ntiles) load_slots(tile + 1);             // the next tile's slots: a global load per tile stalled every warp here
      if (threadIdx.x == 0 && tile + 1 < ntiles) {             // next tile's keys and values into the other buffers
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(bars + (kb ^ 1), kBlobBytes);
        bulk_load(sK_u + (kb ^ 1) * kTileBytes, kblob + static_cast<size_t>(tile + 1) * 2 * kBlobBytes, kBlobBytes, bars + (kb ^ 1));
        mbar_arrive_expect_tx(bars + 2 + (kb ^ 1), kBlobBytes);
        bulk_load(sV_u + (kb ^ 1) * kTileBytes, vblob + static_cast<size_t>(tile + 1) * 2 * kBlobBytes, kBlobBytes, bars + 2 + (kb ^ 1));
      }
      mbar_wait(bars + kb, (phases >> kb) & 1u); phases ^= 1u << kb;
    }
    const int nval = tile < 0 ? kS : min(kS, nv - tile * kS);
    const uint32_t sKb = sK_u + kb * kTileBytes, sVb = sV_u + kb * kTileBytes;

#if HMVIT_DENSE_HEADWARP
    {
      // warp == head hd of the head group; hh runs over the four 16-row query blocks.  The K fragments of the tile are
      // loaded once (before the loop), the V fragments once (after the first block's QK^T, when the values have landed).
      const int hd = warp;
      uint32_t kbA[4][4], kbB[4][4], vbAll[4][2][4];
      {
        const int krow = (lane & 7) + (lane >> 4) * 8;
        const int kunit = hd * 4 + ((lane >> 3) & 1);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          ldsm_x4(sKb + tile_off(np * 16 + krow, kunit), kbA[np][0], kbA[np][1], kbA[np][2], kbA[np][3]);
          ldsm_x4(sKb + tile_off(np * 16 + krow, kunit + 2), kbB[np][0], kbB[np][1], kbB[np][2], kbB[np][3]);
        }
      }
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const float* bh = sBias + hd * kBiasStride + (2 * hh + 7) * 15 + (g + 7);
        float sacc[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const uint32_t kf = koff[nt >> 1] >> ((nt & 1) * 16);
          const int k0 = kf & 0xff, k1 = (kf >> 8) & 0xff;
          sacc[nt][0] = bh[-k0];      sacc[nt][1] = bh[-k1];
          sacc[nt][2] = bh[15 - k0];  sacc[nt][3] = bh[15 - k1];
        }
        {
          uint32_t qa[2][4];
          const int qrow = hh * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldsm_x4(sQ_u + tile_off(qrow, hd * 4 + (lane >> 4)), qa[0][0], qa[0][1], qa[0][2], qa[0][3]);
          ldsm_x4(sQ_u + tile_off(qrow, hd * 4 + 2 + (lane >> 4)), qa[1][0], qa[1][1], qa[1][2], qa[1][3]);
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16(sacc[np * 2], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kbA[np][0], kbA[np][1]);
            mma_bf16(sacc[np * 2 + 1], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kbA[np][2], kbA[np][3]);
          }
#pragma unroll
          for (int np = 0; np < 4; ++np) {
            mma_bf16(sacc[np * 2], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kbB[np][0], kbB[np][1]);
            mma_bf16(sacc[np * 2 + 1], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kbB[np][2], kbB[np][3]);
          }
        }
        if (nval < kS) {                                          // tail of the last tile
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (nt * 8 + 2 * t + e >= nval) { sacc[nt][e] = -INFINITY; sacc[nt][2 + e] = -INFINITY; }
            }
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
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
        uint32_t pa[4][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float p0 = ex2(sacc[nt][0] - mu0), p1 = ex2(sacc[nt][1] - mu0);
          const float p2 = ex2(sacc[nt][2] - mu1), p3 = ex2(sacc[nt][3] - mu1);
          pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
          pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
        float lacc[4] = {lrow[hh][0] * al0, 0.f, lrow[hh][1] * al1, 0.f};
#pragma unroll
        for (int n = 0; n < 4; ++n) { o[hh][n][0] *= al0; o[hh][n][1] *= al0; o[hh][n][2] *= al1; o[hh][n][3] *= al1; }
        if (hh == 0) {
          if (tile >= 0) { mbar_wait(bars + 2 + kb, (phases >> (2 + kb)) & 1u); phases ^= 4u << kb; }     // this tile's values have landed
          const int vrow = (lane & 7) + ((lane >> 3) & 1) * 8;
          const int vunit = hd * 4 + (lane >> 4);
#pragma unroll
          for (int kc = 0; kc < 4; ++kc)
#pragma unroll
            for (int np = 0; np < 2; ++np)
              ldsm_x4_t(sVb + tile_off(kc * 16 + vrow, vunit + np * 2), vbAll[kc][np][0], vbAll[kc][np][1], vbAll[kc][np][2], vbAll[kc][np][3]);
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          mma_bf16(lacc, pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], 0x3F803F80u, 0x3F803F80u);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            mma_bf16(o[hh][np * 2], pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], vbAll[kc][np][0], vbAll[kc][np][1]);
            mma_bf16(o[hh][np * 2 + 1], pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], vbAll[kc][np][2], vbAll[kc][np][3]);
          }
        }
        lrow[hh][0] = lacc[0]; lrow[hh][1] = lacc[2];
      }
    }
#else
#pragma unroll
    for (int hh = 0; hh < kHPW; ++hh) {
      const int hd = h0 + hh;                                   // head within the head group
      const float* bh = sBias + hd * kBiasStride + bias_q;
      // (adding the bias after the MMAs instead of starting the accumulators from it, so that the tensor-core chain
      //  does not wait on shared memory, was measured slower: 0.382 vs 0.368 ms)
      float sacc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t kf = koff[nt >> 1] >> ((nt & 1) * 16);
        const int k0 = kf & 0xff, k1 = (kf >> 8) & 0xff;
        sacc[nt][0] = bh[-k0];      sacc[nt][1] = bh[-k1];
        sacc[nt][2] = bh[15 - k0];  sacc[nt][3] = bh[15 - k1];
      }
      {
        // S = Q_h K_h^T: every operand fragment is requested before the MMAs that consume it are issued (the
        // asm statements keep their order, so the loads of the second k-step are interleaved by hand)
        uint32_t qa[2][4], kbA[4][4], kbB[4][4];
        const int qrow = rb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int krow = (lane & 7) + (lane >> 4) * 8;
        const int kunit = hd * 4 + ((lane >> 3) & 1);
        ldsm_x4(sQ_u + tile_off(qrow, hd * 4 + (lane >> 4)), qa[0][0], qa[0][1], qa[0][2], qa[0][3]);
        ldsm_x4(sQ_u + tile_off(qrow, hd * 4 + 2 + (lane >> 4)), qa[1][0], qa[1][1], qa[1][2], qa[1][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) ldsm_x4(sKb + tile_off(np * 16 + krow, kunit), kbA[np][0], kbA[np][1], kbA[np][2], kbA[np][3]);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          ldsm_x4(sKb + tile_off(np * 16 + krow, kunit + 2), kbB[np][0], kbB[np][1], kbB[np][2], kbB[np][3]);
          mma_bf16(sacc[np * 2], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kbA[np][0], kbA[np][1]);
          mma_bf16(sacc[np * 2 + 1], qa[0][0], qa[0][1], qa[0][2], qa[0][3], kbA[np][2], kbA[np][3]);
        }
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          mma_bf16(sacc[np * 2], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kbB[np][0], kbB[np][1]);
          mma_bf16(sacc[np * 2 + 1], qa[1][0], qa[1][1], qa[1][2], qa[1][3], kbB[np][2], kbB[np][3]);
        }
      }
      if (nval < kS) {                                          // tail of the last tile
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (nt * 8 + 2 * t + e >= nval) { sacc[nt][e] = -INFINITY; sacc[nt][2 + e] = -INFINITY; }
          }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
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
      // running denominator from exactly the bf16 probabilities that multiply V (P x ones on the tensor core);
      // its four MMAs ride inside the PV loop instead of forming a dependent chain of their own
      float lacc[4] = {lrow[hh][0] * al0, 0.f, lrow[hh][1] * al1, 0.f};
#pragma unroll
      for (int n = 0; n < 4; ++n) { o[hh][n][0] *= al0; o[hh][n][1] *= al0; o[hh][n][2] *= al1; o[hh][n][3] *= al1; }
      if (hh == 0 && tile >= 0) { mbar_wait(bars + 2 + kb, (phases >> (2 + kb)) & 1u); phases ^= 4u << kb; }     // this tile's values have landed
      {
        uint32_t vb[2][2][4];                                    // [buffer][dim n-tile pair][fragment]
        const int vrow = (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vunit = hd * 4 + (lane >> 4);
#pragma unroll
        for (int np = 0; np < 2; ++np) ldsm_x4_t(sVb + tile_off(vrow, vunit + np * 2), vb[0][np][0], vb[0][np][1], vb[0][np][2], vb[0][np][3]);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          const int cur = kc & 1;
          if (kc < 3) {
#pragma unroll
            for (int np = 0; np < 2; ++np)
              ldsm_x4_t(sVb + tile_off((kc + 1) * 16 + vrow, vunit + np * 2), vb[cur ^ 1][np][0], vb[cur ^ 1][np][1], vb[cur ^ 1][np][2], vb[cur ^ 1][np][3]);
          }
          mma_bf16(lacc, pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], 0x3F803F80u, 0x3F803F80u);
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            mma_bf16(o[hh][np * 2], pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], vb[cur][np][0], vb[cur][np][1]);
            mma_bf16(o[hh][np * 2 + 1], pa[kc][0], pa[kc][1], pa[kc][2], pa[kc][3], vb[cur][np][2], vb[cur][np][3]);
          }
        }
      }
      lrow[hh][0] = lacc[0]; lrow[hh][1] = lacc[2];
    }
#endif
    __syncthreads();                                           // every warp is done with this tile's K and V buffers
  }

  // ---- training: save the softmax statistics (log2 domain: running max + log2 of the denominator) ----
  if (p.lse != nullptr && t == 0) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
#if HMVIT_DENSE_HEADWARP
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        int r, c; group_token(p.kind, gy, gx, hh * 16 + g + e * 8, p.H, p.W, r, c);
        p.lse[(static_cast<size_t>(a) * N + r * p.W + c) * kHeads + hgc * kHG + warp] = lrow[hh][e] > 0.f ? mrow[hh][e] + log2f(lrow[hh][e]) : INFINITY;
      }
#else
      int r, c; group_token(p.kind, gy, gx, rb * 16 + g + e * 8, p.H, p.W, r, c);
      float* dst = p.lse + (static_cast<size_t>(a) * N + r * p.W + c) * kHeads + hgc * kHG + h0;
#pragma unroll
      for (int hh = 0; hh < kHPW; ++hh) dst[hh] = lrow[hh][e] > 0.f ? mrow[hh][e] + log2f(lrow[hh][e]) : INFINITY;
#endif
    }
  }
  // ---- normalise, stage in smem (K buffer 0: no copy is pending), coalesced store ----
#pragma unroll
  for (int hh = 0; hh < kHPW; ++hh) {
    const float il0 = lrow[hh][0] > 0.f ? 1.0f / lrow[hh][0] : 0.f;
    const float il1 = lrow[hh][1] > 0.f ? 1.0f / lrow[hh][1] : 0.f;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
#if HMVIT_DENSE_HEADWARP
      const int col = warp * kDh + n * 8 + 2 * t;
      const int r0 = hh * 16 + g, r1 = r0 + 8;
#else
      const int col = (h0 + hh) * kDh + n * 8 + 2 * t;
      const int r0 = rb * 16 + g, r1 = r0 + 8;
#endif
      *reinterpret_cast<uint32_t*>(sK + tile_off(r0, col >> 3) + (col & 7) * 2) = pack_bf16x2(o[hh][n][0] * il0, o[hh][n][1] * il0);
      *reinterpret_cast<uint32_t*>(sK + tile_off(r1, col >> 3) + (col & 7) * 2) = pack_bf16x2(o[hh][n][2] * il1, o[hh][n][3] * il1);
    }
  }
  __syncthreads();
  {
    uint4* dst = reinterpret_cast<uint4*>(p.out) + static_cast<size_t>(a) * N * 32 + cu0 + u16;
#pragma unroll
    for (int tt = 0; tt < kStageIt; ++tt) {
      const int s = warp * (2 * kStageIt) + tt * 2 + hl;
      int r, c; group_token(p.kind, gy, gx, s, p.H, p.W, r, c);
      dst[static_cast<size_t>(r * p.W + c) * 32] = *reinterpret_cast<const uint4*>(sK + tile_off(s, u16));
    }
  }
}

}  // namespace hmvit
