// Fused warp + mask + multi-agent group attention, persistent and warp-specialised, on the 5th-gen tensor
// cores (tcgen05 + TMEM).  Default implementation of hmvit_group_attn; replaces
//   HeteroFusionBlock.warp_features            hetero_fusion.py:338-361
//   the ego loop around HeteroAttention.forward hetero_fusion.py:373-397 / 412-440, 187-277
//   get_roi_and_cav_mask / warp_affine          torch_transformation_utils.py:11-134, 254-355
//
// Two kernels:
//
//   tap_records_kernel   once per partition kind per FORWARD (the poses do not change between the block
//                        iterations): for every (scene b, ego i, token group g) the fp64 source-pixel map of every
//                        (source j, token) of the group, the bit-exact ROI visibility, and the COMPACTED list of
//                        visible keys as 16-byte records (tap corner, 4 bilinear weights as bf16, source, slot,
//                        relative-position offset).  The ego's own keys are ordinary records (identity pose ->
//                        one tap of weight 1).
//
//   fused_attn_kernel    persistent CTAs (2 per SM), each owning one head group (4 heads = 2 head PAIRS) and looping
//                        over work items (b, i, g).  The blended key / value tiles never touch HBM:
//     warps 4-11  GATHER   half-warp == key: the record's (up to) 4 tap rows of the projected K' / V' planes
//                          (16-byte loads, K and V batches software-pipelined), packed-bf16 blend on top of the
//                          folded bias, rows stored as UMMA SWIZZLE_128B operand tiles into a 2-stage ring
//     warps 13-15 Q LOAD   the item's 64 query rows into the block-diagonal Q tiles (zero halves are static)
//     warp  12    MMA      converged warp, one elected lane:  S = Qbd K^T  (M128 N64 K64: two heads stacked on M),
//                          D += P V  (P from TMEM, V MN-major)
//     warps 0-3   SOFTMAX  thread == TMEM lane == (head of the pair, query row); the two pairs ping-pong: S from
//                          TMEM, + relative position bias (through the tile's key-offset list), running max,
//                          exp2, bf16 P over S in TMEM, D rescaled in TMEM only when the max moved; normalise + store
//   Work items of one ego are consecutive, so its sources' K' / V' planes stay L2-resident; CTA prologue
//   (TMEM allocation, barrier init, bias tables, static zeros) is paid once per launch, not once per item.
#pragma once
#include "attn_split.cuh"

namespace hmvit {

// MN-major operand (rows of 128 B = 64 consecutive MN elements for one k; 8-row groups 1024 B apart)
HMVIT_DEVINL uint64_t umma_desc_sw128_mn(uint32_t smem_addr) { return umma_desc_sw128(smem_addr); }

// one visible key of a (scene, ego, group): 16 bytes
struct KeyRec {
  short x0, y0;                // (y0, x0) corner of the bilinear footprint in the source map
  uint32_t w01, w23;           // tap weights as bf16 pairs: (w00, w01), (w10, w11); 0 = tap outside the map
  uint32_t meta;               // source j | slot << 8 | source type << 16 | relative-position offset << 24
};
static_assert(sizeof(KeyRec) == 16, "KeyRec must be 16 bytes");

constexpr int kRecMaxL = 8;    // agents per scene the record pass handles (2 x 256 entries per CTA)
constexpr int kFusedMaxAgents = 1024;   // B * L the persistent kernel's item table holds

struct RecParams {
  AttnParams a;                // geometry: B, L, H, W, mode, record_len, cav_mask, T, cell, key_mask
  KeyRec* rec;                 // [nkinds][B*L][G][L*64]
  int* nvis;                   // [nkinds][B*L][G] visible keys of the (ego, group)
  int kind0;                   // kind of grid.z == 0
};

// grid (G, B*L, nkinds)
__global__ void __launch_bounds__(256, 4) tap_records_kernel(const RecParams rp) {
  const AttnParams& p = rp.a;
  const int a = blockIdx.y;
  const int b = a / p.L, i = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);
  if (i >= nrec) return;
  const int kind = rp.kind0 + blockIdx.z;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int grp = blockIdx.x;
  const int gy = grp / GX, gx = grp - gy * GX;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  __shared__ int sCnt[kRecMaxL * 2];
  const int nent = nrec * kS;
  KeyRec rec[2];
  uint32_t bal[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int e = it * 256 + threadIdx.x;
    KeyRec r; r.x0 = 0; r.y0 = 0; r.w01 = 0; r.w23 = 0; r.meta = 0;
    bool vis = false;
    if (e < nent) {
      const int j = e >> 6, tk = e & 63;
      const int tj = p.mode[b * p.L + j] != 0 ? 1 : 0;
      r.meta = static_cast<uint32_t>(j) | (static_cast<uint32_t>(tk) << 8) | (static_cast<uint32_t>(tj) << 16) |
               (static_cast<uint32_t>((tk >> 3) * 15 + (tk & 7)) << 24);
      if (p.cav_mask[b * p.L + j] != 0) {
        const WarpMap wm = make_warp_map(p.T + ((static_cast<size_t>(b) * p.L + j) * p.L + i) * 16, p.H, p.W, p.cell);
        int rr, cc; group_token(kind, gy, gx, tk, p.H, p.W, rr, cc);
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
  const size_t ag = (static_cast<size_t>(blockIdx.z) * p.B * p.L + a) * G + grp;
  KeyRec* dst = rp.rec + ag * (static_cast<size_t>(p.L) * kS);
  int nv = 0;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int seg = it * 8 + warp;
    int base = 0;
#pragma unroll
    for (int s = 0; s < 2 * kRecMaxL; ++s) {
      const int c = sCnt[s];
      if (s < seg) base += c;
      nv += (it == 0) ? c : 0;
    }
    if ((bal[it] >> lane) & 1u)
      *reinterpret_cast<uint4*>(dst + base + __popc(bal[it] & ((1u << lane) - 1u))) = *reinterpret_cast<const uint4*>(&rec[it]);
  }
  if (threadIdx.x == 0) rp.nvis[ag] = nv;
}

// ------------------------------------------------------------------------------------------------
struct FusedAttnParams {
  AttnParams a;
  const KeyRec* rec;           // [B*L][G][L*64] records of a.kind
  const int* nvis;             // [B*L][G]
};

#ifndef HMVIT_FA_SOFT_REGS   // registers per thread after rebalancing: softmax warpgroup / gather warpgroups / MMA + Q-load warpgroup
#define HMVIT_FA_SOFT_REGS 112
#endif
#ifndef HMVIT_FA_GATHER_REGS
#define HMVIT_FA_GATHER_REGS 56
#endif
#ifndef HMVIT_FA_MISC_REGS
#define HMVIT_FA_MISC_REGS 32
#endif
#ifndef HMVIT_FA_DBG         // bottleneck-hunting builds only (results are wrong): 1 no tap loads, 2 no softmax math, 4 no MMAs
#define HMVIT_FA_DBG 0
#endif

struct FaCfg {
  static constexpr int THREADS = 512;                   // 4 warpgroups: softmax | gather | gather | MMA + Q load
  static constexpr int GATHER_WARPS = 8;
  static constexpr int QLOAD_THREADS = 96;
  static constexpr int OFF_Q = 0;                       // [2 pairs][128 rows][128 B]  block-diagonal Q
  static constexpr int OFF_KV = 32768;                  // [2 stages][K: 2 pairs x 8 KB | V: 2 pairs x 8 KB]
  static constexpr int KV_STAGE = 32768;
  static constexpr int OFF_BIAS = OFF_KV + 2 * KV_STAGE;               // [4 heads][kBiasStride] fp32, log2 domain
  static constexpr int OFF_KVB = OFF_BIAS + kHG * kBiasStride * 4;     // [2 te][2 tj][K | V][128 ch] bf16 folded biases
  static constexpr int OFF_KOFF = OFF_KVB + 2 * 2 * 2 * 256;           // [4 tiles][64] byte offset of every key's bias column
  static constexpr int OFF_VALID = OFF_KOFF + 4 * 64 * 4;              // [kFusedMaxAgents] uint16: agents that own work items
  static constexpr int OFF_BAR = OFF_VALID + kFusedMaxAgents * 2;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;              // + alignment slack
  static constexpr uint32_t TM_COLS = 256;              // pair p: S / P at 128 p, D at 128 p + 64
};

HMVIT_DEVINL float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__global__ void __launch_bounds__(FaCfg::THREADS, 2) fused_attn_kernel(const FusedAttnParams fp) {
  using Cfg = FaCfg;
  const AttnParams& p = fp.a;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sQ = smem + Cfg::OFF_Q;
  uint8_t* sKV = smem + Cfg::OFF_KV;
  float* sBias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint8_t* sKvb = smem + Cfg::OFF_KVB;
  int* sKoff = reinterpret_cast<int*>(smem + Cfg::OFF_KOFF);
  uint16_t* sValid = reinterpret_cast<uint16_t*>(smem + Cfg::OFF_VALID);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars + 0;      // [2 stages]  gather warps -> MMA
  uint64_t* kv_empty = bars + 2;     // [2 stages]  MMA (commit) -> gather warps
  uint64_t* s_full = bars + 4;       // [2 pairs]   MMA (commit) -> softmax
  uint64_t* p_full = bars + 6;       // [2 pairs]   softmax -> MMA
  uint64_t* d_full = bars + 8;       // [2 pairs]   MMA (commit) -> softmax: the item's accumulator is final
  uint64_t* d_free = bars + 10;      // [2 pairs]   softmax -> MMA: accumulator read, the next item may overwrite it
  uint64_t* q_full = bars + 12;      //             Q loaders -> MMA
  uint64_t* q_empty = bars + 13;     //             MMA (commit) -> Q loaders
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  int* s_nvalid = reinterpret_cast<int*>(bars + 15);

  // ------------------------------ one-off prologue ------------------------------
  const int hgc = blockIdx.x & 1;                       // head group of this CTA (items alternate over CTAs)
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], Cfg::GATHER_WARPS); mbar_init(&kv_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128);
      mbar_init(&d_full[s], 1); mbar_init(&d_free[s], 128);
    }
    mbar_init(q_full, Cfg::QLOAD_THREADS); mbar_init(q_empty, 1);
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<Cfg::TM_COLS>(tmem_slot);
  if (warp == 0) {
    // agents that own work items (valid egos; only slot 0 in the dead-query stage), in order
    const int BL = p.B * p.L;
    int n = 0;
    for (int a0 = 0; a0 < BL; a0 += 32) {
      const int a = a0 + lane;
      bool ok = false;
      if (a < BL) {
        const int b = a / p.L, i = a - b * p.L;
        ok = i < min(p.record_len[b], p.L) && !(p.ego_only && i != 0);
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, ok);
      if (ok) sValid[n + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint16_t>(a);
      n += __popc(bal);
    }
    if (lane == 0) *s_nvalid = n;
  }
  // relative position bias of this head group, log2 domain: sBias[h][idx]
  for (int e = tid; e < 225 * kHG; e += Cfg::THREADS)
    sBias[(e & 3) * kBiasStride + (e >> 2)] = __ldg(p.bias_table + (e >> 2) * kHeads + hgc * kHG + (e & 3)) * 1.4426950408889634f;
  // folded key / value biases of this head group as bf16 rows: [(te, tj)][K | V][128 channels]
  for (int e = tid; e < 4 * 2 * 64; e += Cfg::THREADS) {
    const int tt = e >> 7, kv = (e >> 6) & 1, c2 = e & 63;
    const float2 v = __ldg(reinterpret_cast<const float2*>((kv == 0 ? p.bk : p.bv) + tt * kC + hgc * 128) + c2);
    reinterpret_cast<uint32_t*>(sKvb)[e] = pack_bf16x2(v.x, v.y);
  }
  // static zero halves of the block-diagonal Q tiles: pair pr, rows [0,64) hold head 2pr in K-columns [0,32),
  // rows [64,128) hold head 2pr+1 in K-columns [32,64)
  for (int e = tid; e < 2 * 128 * 4; e += Cfg::THREADS) {
    const int pr = e >> 9, row = (e >> 2) & 127, cq = e & 3;
    const int unit = (row < 64 ? 4 : 0) + cq;
    *reinterpret_cast<uint4*>(sQ + pr * 16384 + sw128_offset(row, unit)) = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const int nvalid = *s_nvalid;
  // this CTA's items: (valid agent, group) pairs, strided over the CTAs that share its head group
  const int n_items = nvalid * G;
  const int item0 = blockIdx.x >> 1, item_step = gridDim.x >> 1;

  if (warp < 4) {
    // =========================================== SOFTMAX ===========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HMVIT_FA_SOFT_REGS));
    const int hh = tid >> 6, row = tid & 63;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    // bias of (query row, key slot s') = table[bias_q - koff(s')], koff = (s' >> 3) * 15 + (s' & 7)
    const int bias_q = ((row >> 3) + 7) * 15 + (row & 7) + 7;
    uint32_t tcnt = 0, icnt = 0;
    for (int it = item0; it < n_items; it += item_step) {
      const int a = sValid[it / G], grp = it - (it / G) * G;
      const int gy = grp / GX, gx = grp - gy * GX;
      const int nv = __ldg(fp.nvis + static_cast<size_t>(a) * G + grp);
      const int ntiles = (nv + kS - 1) >> 6;
      int r, c; group_token(p.kind, gy, gx, row, p.H, p.W, r, c);
      const size_t tok = static_cast<size_t>(a) * N + r * p.W + c;
      float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
      for (int t = 0; t < ntiles; ++t, ++tcnt) {
        const int nval = min(kS, nv - t * kS);
        const uint32_t ko_u = smem_u32(sKoff + (tcnt & 3u) * kS);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          const uint32_t tS = tm + lane_base + pr * 128, tD = tS + 64 + hh * 32;
          const uint32_t bt_u = smem_u32(sBias + (pr * 2 + hh) * kBiasStride + bias_q);
          mbar_wait(&s_full[pr], tcnt & 1u);
          tc_fence_after();
          uint32_t sa[32], sb[32];
          tmem_ld32(tS, sa);
          tmem_ld32(tS + 32, sb);
          tmem_ld_wait();
          float mx = -INFINITY;
          if (!(HMVIT_FA_DBG & 2)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 k4 = lds_u4_addr(ko_u + q * 16);
              sa[q * 4 + 0] = __float_as_uint(__uint_as_float(sa[q * 4 + 0]) + lds_f32(bt_u - k4.x));
              sa[q * 4 + 1] = __float_as_uint(__uint_as_float(sa[q * 4 + 1]) + lds_f32(bt_u - k4.y));
              sa[q * 4 + 2] = __float_as_uint(__uint_as_float(sa[q * 4 + 2]) + lds_f32(bt_u - k4.z));
              sa[q * 4 + 3] = __float_as_uint(__uint_as_float(sa[q * 4 + 3]) + lds_f32(bt_u - k4.w));
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 k4 = lds_u4_addr(ko_u + 128 + q * 16);
              sb[q * 4 + 0] = __float_as_uint(__uint_as_float(sb[q * 4 + 0]) + lds_f32(bt_u - k4.x));
              sb[q * 4 + 1] = __float_as_uint(__uint_as_float(sb[q * 4 + 1]) + lds_f32(bt_u - k4.y));
              sb[q * 4 + 2] = __float_as_uint(__uint_as_float(sb[q * 4 + 2]) + lds_f32(bt_u - k4.z));
              sb[q * 4 + 3] = __float_as_uint(__uint_as_float(sb[q * 4 + 3]) + lds_f32(bt_u - k4.w));
            }
            if (nval < kS) {                                     // tail of the item's last tile
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                if (e >= nval) sa[e] = 0xff800000u;              // -inf
                if (32 + e >= nval) sb[e] = 0xff800000u;
              }
            }
#pragma unroll
            for (int e = 0; e < 32; e += 2) {
              mx = fmaxf(mx, fmaxf(__uint_as_float(sa[e]), __uint_as_float(sa[e + 1])));
              mx = fmaxf(mx, fmaxf(__uint_as_float(sb[e]), __uint_as_float(sb[e + 1])));
            }
          } else {
            mx = 0.f;
          }
          const float m_new = fmaxf(m_run[pr], mx);
          const float mu = (m_new == -INFINITY) ? 0.f : m_new;
          const float alpha = ex2(m_run[pr] - mu);               // 0 for the first tile (m_run = -inf)
          m_run[pr] = m_new;
          float ls0 = 0.f, ls1 = 0.f;
          uint32_t pk[32];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            if (HMVIT_FA_DBG & 2) { pk[k] = sa[k]; pk[16 + k] = sb[k]; continue; }
            const float p0 = ex2(__uint_as_float(sa[2 * k]) - mu), p1 = ex2(__uint_as_float(sa[2 * k + 1]) - mu);
            const float p2 = ex2(__uint_as_float(sb[2 * k]) - mu), p3 = ex2(__uint_as_float(sb[2 * k + 1]) - mu);
            ls0 += p0 + p1; ls1 += p2 + p3;
            pk[k] = pack_bf16x2(p0, p1);
            pk[16 + k] = pack_bf16x2(p2, p3);
          }
          l_run[pr] = l_run[pr] * alpha + (ls0 + ls1);
          if (t > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) {
            // the running max moved: rescale this head's accumulator in TMEM (P V of the previous tile has retired:
            // the commit behind s_full covers every earlier MMA)
            uint32_t d[32];
            tmem_ld32(tD, d);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) d[k] = __float_as_uint(__uint_as_float(d[k]) * alpha);
            tmem_st32(tD, d);
          }
          tmem_st32(tS, pk);                                     // P (bf16 pairs, A operand of P V) over the consumed logits
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[pr]);
        }
      }
      // ------------------------------ normalise and store ------------------------------
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const uint32_t tD = tm + lane_base + pr * 128 + 64 + hh * 32;
        const int head = hgc * kHG + pr * 2 + hh;
        uint32_t d[32];
        float il = 0.f;
        if (ntiles > 0) {
          mbar_wait(&d_full[pr], icnt & 1u);
          tc_fence_after();
          tmem_ld32(tD, d);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(&d_free[pr]);                              // accumulator in registers: the next item may overwrite it
          il = l_run[pr] > 0.f ? 1.0f / l_run[pr] : 0.f;
        }
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
        if (p.lse != nullptr) p.lse[tok * kHeads + head] = l_run[pr] > 0.f ? m_run[pr] + log2f(l_run[pr]) : INFINITY;
      }
      if (ntiles > 0) ++icnt;
    }
  } else if (warp < 12) {
    // =========================================== GATHER ===========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HMVIT_FA_GATHER_REGS));
    const int hw = (warp - 4) * 2 + (lane >> 4);        // half-warp 0..15: keys hw, hw + 16, hw + 32, hw + 48 of a tile
    const int u16 = lane & 15, pr = u16 >> 3, un = u16 & 7;
    const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;            // uint4 units per ego-type plane
    const uint32_t kv_u = smem_u32(sKV), kvb_u = smem_u32(sKvb);
    uint32_t tcnt = 0;
    for (int it = item0; it < n_items; it += item_step) {
      const int a = sValid[it / G], grp = it - (it / G) * G;
      const int b = a / p.L;
      const int nv = __ldg(fp.nvis + static_cast<size_t>(a) * G + grp);
      const int ntiles = (nv + kS - 1) >> 6;
      const int te = p.mode[a] != 0 ? 1 : 0;
      const uint4* kbase = reinterpret_cast<const uint4*>(p.k) + te * plane + static_cast<size_t>(b) * p.L * N * 32 + hgc * 16 + u16;
      const uint4* vbase = reinterpret_cast<const uint4*>(p.v) + te * plane + static_cast<size_t>(b) * p.L * N * 32 + hgc * 16 + u16;
      const uint4* recs = reinterpret_cast<const uint4*>(fp.rec + (static_cast<size_t>(a) * G + grp) * (static_cast<size_t>(p.L) * kS));
      const uint32_t kvb_te = kvb_u + te * 1024 + u16 * 16;
      for (int t = 0; t < ntiles; ++t, ++tcnt) {
        const int nval = min(kS, nv - t * kS);
        const uint32_t stage = tcnt & 1u;
        // this half-warp's 4 records of the tile: lane q of the half-warp fetches record q, shuffled out below
        uint4 myrec = make_uint4(0, 0, 0, 0);
        {
          const int key = hw + (u16 & 3) * 16;
          if (key < nval) myrec = __ldg(recs + t * kS + key);
        }
        mbar_wait(&kv_empty[stage], ((tcnt >> 1) & 1u) ^ 1u);      // P V of the tile two back has retired
        const uint32_t dstK = kv_u + stage * Cfg::KV_STAGE + pr * 8192, dstV = dstK + 16384;
        int* koff = sKoff + (tcnt & 3u) * kS;

        uint4 kk[4], vv[4];
        uint32_t wq[4], meta = 0;
        auto fetch = [&](int n) {                                  // record n of this half-warp -> tap weights
          const int src = (lane & 16) | n;
          const uint32_t r0 = __shfl_sync(0xffffffffu, myrec.x, src), r1 = __shfl_sync(0xffffffffu, myrec.y, src);
          const uint32_t r2 = __shfl_sync(0xffffffffu, myrec.z, src), r3 = __shfl_sync(0xffffffffu, myrec.w, src);
          wq[0] = r1 & 0xffffu; wq[1] = r1 >> 16; wq[2] = r2 & 0xffffu; wq[3] = r2 >> 16;
          meta = r3;
          const int x0 = static_cast<short>(r0 & 0xffffu), y0 = static_cast<short>(r0 >> 16);
          return static_cast<int>(((r3 & 0xffu) * N + y0 * p.W + x0) * 32);   // uint4 offset of the (y0, x0) row in the scene's planes
        };
        auto issue = [&](const uint4* base, int off, uint4 (&tv)[4]) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            tv[q] = make_uint4(0, 0, 0, 0);
            if (wq[q] != 0u && !(HMVIT_FA_DBG & 1))                // a tap outside the map has weight 0 and is not loaded
              tv[q] = __ldg(base + off + ((q >> 1) * p.W + (q & 1)) * 32);
          }
        };
        auto blend_store = [&](const uint4 (&tv)[4], uint32_t bias_addr, uint32_t dst, int key) {
          uint4 o = make_uint4(0, 0, 0, 0);
          if (key < nval) {
            o = lds_u4_addr(bias_addr);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t w2 = wq[q] | (wq[q] << 16);
              o.x = hfma2_bf16(w2, tv[q].x, o.x); o.y = hfma2_bf16(w2, tv[q].y, o.y);
              o.z = hfma2_bf16(w2, tv[q].z, o.z); o.w = hfma2_bf16(w2, tv[q].w, o.w);
            }
          }
          sts_u4_addr(dst + sw128_offset(key, un), o);
        };
        // software pipeline over the 4 keys: the value taps of key n are in flight while its key taps are blended,
        // the key taps of key n + 1 while its value taps are blended
        int off = fetch(0);
        issue(kbase, off, kk);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int key = hw + n * 16;
          const uint32_t tj = (meta >> 16) & 1u;
          const uint32_t wcur[4] = {wq[0], wq[1], wq[2], wq[3]};
          issue(vbase, off, vv);
          blend_store(kk, kvb_te + tj * 512, dstK, key);
          if (u16 == 0) koff[key] = key < nval ? static_cast<int>((meta >> 24) * 4u) : 0;
          if (n + 1 < 4) { off = fetch(n + 1); issue(kbase, off, kk); }
          {
            // blend the values with the weights of key n (wq now holds key n + 1's)
            uint4 o = make_uint4(0, 0, 0, 0);
            if (key < nval) {
              o = lds_u4_addr(kvb_te + tj * 512 + 256);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t w2 = wcur[q] | (wcur[q] << 16);
                o.x = hfma2_bf16(w2, vv[q].x, o.x); o.y = hfma2_bf16(w2, vv[q].y, o.y);
                o.z = hfma2_bf16(w2, vv[q].z, o.z); o.w = hfma2_bf16(w2, vv[q].w, o.w);
              }
            }
            sts_u4_addr(dstV + sw128_offset(key, un), o);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_full[stage]);
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HMVIT_FA_MISC_REGS));   // one instruction for the whole warpgroup (warps 12-15)
  if (warp == 12) {
    // =========================================== MMA ===========================================
    constexpr uint32_t idesc_qk = umma_idesc(1u, 128, 64);
    constexpr uint32_t idesc_pv = umma_idesc(1u, 128, 64) | (1u << 16);        // B (V tile) MN-major
    const uint32_t q_u = smem_u32(sQ), kv_u = smem_u32(sKV);
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    auto issue_qk = [&](int pr, uint32_t stage) {                 // S_pr = Qbd_pr K_pr^T
      if (HMVIT_FA_DBG & 4) return;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_ss<2>(tmu + pr * 128, umma_desc_sw128(q_u + pr * 16384 + ks * 32),
                   umma_desc_sw128(kv_u + stage * Cfg::KV_STAGE + pr * 8192 + ks * 32), idesc_qk, ks != 0 ? 1u : 0u);
    };
    auto issue_pv = [&](int pr, uint32_t stage, bool first) {     // D_pr (+)= P_pr V_pr
      if (HMVIT_FA_DBG & 4) return;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_ts_bf16(tmu + pr * 128 + 64, tmu + pr * 128 + ks * 8,
                     umma_desc_sw128_mn(kv_u + stage * Cfg::KV_STAGE + 16384 + pr * 8192 + ks * 2048), idesc_pv,
                     (!first || ks != 0) ? 1u : 0u);
    };
    uint32_t tcnt = 0, icnt = 0;
    for (int it = item0; it < n_items; it += item_step) {
      const int a = sValid[it / G], grp = it - (it / G) * G;
      const int nv = __ldg(fp.nvis + static_cast<size_t>(a) * G + grp);
      const int ntiles = (nv + kS - 1) >> 6;
      if (ntiles == 0) continue;
      mbar_wait(q_full, icnt & 1u);
      mbar_wait(&kv_full[tcnt & 1u], (tcnt >> 1) & 1u);
      tc_fence_after();
      if (elect_one()) {
        issue_qk(0, tcnt & 1u); umma_commit(&s_full[0]);
        issue_qk(1, tcnt & 1u); umma_commit(&s_full[1]);
        if (ntiles == 1) umma_commit(q_empty);                    // the item's last Q K^T has been issued
      }
      __syncwarp();
      for (int t = 0; t < ntiles; ++t, ++tcnt) {
        const uint32_t stage = tcnt & 1u;
        const bool has_next = t + 1 < ntiles;
        if (has_next) { mbar_wait(&kv_full[stage ^ 1u], ((tcnt + 1) >> 1) & 1u); tc_fence_after(); }
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          mbar_wait(&p_full[pr], tcnt & 1u);
          if (t == 0) mbar_wait(&d_free[pr], (icnt & 1u) ^ 1u);   // the previous item's accumulator has been read
          tc_fence_after();
          if (elect_one()) {
            issue_pv(pr, stage, t == 0);
            if (has_next) { issue_qk(pr, stage ^ 1u); umma_commit(&s_full[pr]); }
            else umma_commit(&d_full[pr]);
            if (pr == 1) {
              umma_commit(&kv_empty[stage]);                      // both pairs' P V of this tile issued
              if (has_next && t + 2 == ntiles) umma_commit(q_empty);
            }
          }
          __syncwarp();
        }
      }
      ++icnt;
    }
  } else {
    // =========================================== Q LOAD ===========================================
    const int e0 = tid - 13 * 32;                                  // 0..95
    uint32_t icnt = 0;
    for (int it = item0; it < n_items; it += item_step) {
      const int a = sValid[it / G], grp = it - (it / G) * G;
      const int gy = grp / GX, gx = grp - gy * GX;
      const int nv = __ldg(fp.nvis + static_cast<size_t>(a) * G + grp);
      if (nv == 0) continue;
      const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + hgc * 16;
      mbar_wait(q_empty, (icnt & 1u) ^ 1u);                        // the previous item's last Q K^T has retired
      // 64 tokens x 16 units (256 B of this head group), 4 loads in flight per thread
      for (int e = e0; e < 1024; e += 4 * Cfg::QLOAD_THREADS) {
        uint4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = e + k * Cfg::QLOAD_THREADS;
          if (idx < 1024) {
            int r, c; group_token(p.kind, gy, gx, idx >> 4, p.H, p.W, r, c);
            v[k] = __ldg(qsrc + static_cast<size_t>(r * p.W + c) * 32 + (idx & 15));
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = e + k * Cfg::QLOAD_THREADS;
          if (idx < 1024) {
            const int s = idx >> 4, u = idx & 15;
            const int prq = u >> 3, hq = (u >> 2) & 1, cq = u & 3;
            *reinterpret_cast<uint4*>(sQ + prq * 16384 + sw128_offset(hq * 64 + s, hq * 4 + cq)) = v[k];
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(q_full);
      ++icnt;
    }
  }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TM_COLS>(tm);
  }
}

}  // namespace hmvit
