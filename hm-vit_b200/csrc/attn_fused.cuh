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
//                          + the item's 64 query rows into the block-diagonal Q tiles (zero halves are static)
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
  uint32_t off;                // row offset of the footprint corner in the scene's K' / V' planes, in 16-byte units:
                               // ((source j * N + y0 * W + x0) * 32); the corner lies in the map with both neighbours
  uint32_t w01, w23;           // tap weights as bf16 pairs: (w00, w01), (w10, w11); 0 = tap that fell outside the map
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
    KeyRec r; r.off = 0; r.w01 = 0; r.w23 = 0; r.meta = 0;
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
          Taps tp = make_taps(sx, sy, p.H, p.W);
          // The footprint corner is moved INTO the map (a visible key has floor(sx) in [-1, W-1]): the taps that fell
          // outside carry weight 0 and now read in-map rows, so the gather can load all four taps unconditionally
          // (0 * finite == 0 exactly) -- no per-tap predicates, no zero-initialised destination registers.
          if (tp.x0 < 0) { tp.x0 = 0; tp.w00 = tp.w01; tp.w10 = tp.w11; tp.w01 = 0.f; tp.w11 = 0.f; }
          else if (tp.x0 > p.W - 2) { tp.x0 = p.W - 2; tp.w01 = tp.w00; tp.w11 = tp.w10; tp.w00 = 0.f; tp.w10 = 0.f; }
          if (tp.y0 < 0) { tp.y0 = 0; tp.w00 = tp.w10; tp.w01 = tp.w11; tp.w10 = 0.f; tp.w11 = 0.f; }
          else if (tp.y0 > p.H - 2) { tp.y0 = p.H - 2; tp.w10 = tp.w00; tp.w11 = tp.w01; tp.w00 = 0.f; tp.w01 = 0.f; }
          r.off = static_cast<uint32_t>((j * N + tp.y0 * p.W + tp.x0) * 32);
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
#define HMVIT_FA_SOFT_REGS 120
#endif
#ifndef HMVIT_FA_GATHER_REGS
#define HMVIT_FA_GATHER_REGS 56
#endif
#ifndef HMVIT_FA_MISC_REGS
#define HMVIT_FA_MISC_REGS 24
#endif
#ifndef HMVIT_FA_DBG         // bottleneck-hunting builds only (results are wrong): 1 no tap loads, 2 no softmax math, 4 no MMAs
#define HMVIT_FA_DBG 0
#endif

#ifdef HMVIT_TS   // timeline build (tools/fa_timeline.py): CTAs 0-1 record clock64() per role at pipeline events
__device__ unsigned long long g_fa_ts[2][4][1024];      // [cta][role: 0 softmax, 1 gather, 2 mma, 3 q-load][event]
#define FA_TS(role, idx, code) do { const int i_ = (idx); if (blockIdx.x < 2 && i_ < 1024) g_fa_ts[blockIdx.x][role][i_] = (static_cast<unsigned long long>(clock64()) << 8) | (code); } while (0)
#else
#define FA_TS(role, idx, code) do { } while (0)
#endif

struct FaCfg {
  static constexpr int THREADS = 512;                   // 4 warpgroups: softmax | gather | gather | MMA + Q load
  static constexpr int GATHER_WARPS = 8;
  static constexpr int OFF_Q = 0;                       // [2 pairs][128 rows][128 B]  block-diagonal Q
  static constexpr int OFF_KV = 32768;                  // [2 stages][K: 2 pairs x 8 KB | V: 2 pairs x 8 KB]
  static constexpr int KV_STAGE = 32768;
  static constexpr int OFF_BIAS = OFF_KV + 2 * KV_STAGE;               // [4 heads][kBiasStride] fp32, log2 domain
  static constexpr int OFF_KVB = OFF_BIAS + kHG * kBiasStride * 4;     // [2 te][2 tj][K | V][128 ch] bf16 folded biases
  static constexpr int OFF_KOFF = OFF_KVB + 2 * 2 * 2 * 256;           // [4 tiles][64] byte offset of every key's bias column
  static constexpr int OFF_VALID = OFF_KOFF + 4 * 64 * 4;              // [kFusedMaxAgents] uint16: agents that own work items
  static constexpr int OFF_NV = OFF_VALID + kFusedMaxAgents * 2;       // [NV_ITEMS] uint16: visible keys of this CTA's k-th item
  static constexpr int NV_ITEMS = 2048;
  static constexpr int OFF_BAR = OFF_NV + NV_ITEMS * 2;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;              // + alignment slack
  static constexpr uint32_t TM_COLS = 256;              // pair p: S / P at 128 p, D at 128 p + 64
};

// 16-byte read-only global load, issued in program order; pred == false: no load, zeros
HMVIT_DEVINL uint4 ldg_nc_u4_if(const uint4* p, bool pred) {
  uint4 v;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
      "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p), "r"(pred ? 1 : 0));
  return v;
}
// unconditional 16-byte read-only global load, issued in program order
HMVIT_DEVINL uint4 ldg_nc_u4_v(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
HMVIT_DEVINL uint32_t hfma2_bf16_v(uint32_t a, uint32_t b, uint32_t c) {      // program-order variant of hfma2_bf16
  uint32_t d;
  asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
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
  uint16_t* sNv = reinterpret_cast<uint16_t*>(smem + Cfg::OFF_NV);
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
    mbar_init(q_full, Cfg::GATHER_WARPS); mbar_init(q_empty, 1);
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
  // Visible-key counts of this CTA's items, staged in shared memory once: a per-item global load sat on every role's
  // critical path at each item boundary (a register prefetch does not survive: ptxas spills it right behind the load,
  // which waits for it -- 18 % of the softmax warps' stall samples).  Items beyond the table fall back to global loads.
  for (int k = tid; k < Cfg::NV_ITEMS; k += Cfg::THREADS) {
    const int it_ = item0 + k * item_step;
    if (it_ >= n_items) break;
    sNv[k] = static_cast<uint16_t>(__ldg(fp.nvis + static_cast<size_t>(sValid[it_ / G]) * G + (it_ - (it_ / G) * G)));
  }
  __syncthreads();
  auto item_nv = [&](int k_, int it_) -> int {                    // visible keys of this CTA's k-th item (= item it_)
    if (k_ < Cfg::NV_ITEMS) return sNv[k_];
    return __ldg(fp.nvis + static_cast<size_t>(sValid[it_ / G]) * G + (it_ - (it_ / G) * G));
  };

  if (warp < 4) {
    // =========================================== SOFTMAX ===========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HMVIT_FA_SOFT_REGS));
    const int hh = tid >> 6, row = tid & 63;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    // bias of (query row, key slot s') = table[bias_q - koff(s')], koff = (s' >> 3) * 15 + (s' & 7)
    const int bias_q = ((row >> 3) + 7) * 15 + (row & 7) + 7;
    uint32_t tcnt = 0, icnt = 0;
    int ts_i = 0; (void)ts_i;
    int kit = 0;
    for (int it = item0; it < n_items; it += item_step, ++kit) {
      const int a = sValid[it / G], grp = it - (it / G) * G;
      const int gy = grp / GX, gx = grp - gy * GX;
      const int nv = item_nv(kit, it);
      const int ntiles = (nv + kS - 1) >> 6;
      int r, c; group_token(p.kind, gy, gx, row, p.H, p.W, r, c);
      const size_t tok = static_cast<size_t>(a) * N + r * p.W + c;
      // The softmax code is deliberately COMPACT (loops over pairs / chunks are not unrolled): one warp per SM
      // sub-partition runs it, so nothing amortises instruction fetch -- the fully unrolled form (16 K instructions)
      // ran at IPC 0.1 with "no instruction" as its top stall reason (profiles/r2_attention_study.md).
      // (m_c, l_c) belong to the pair being processed, (m_o, l_o) to the other one; swapped after every step.
      // m: the maximum the exponentials refer to.  It follows the running RAW maximum (no bias) lazily: it only moves
      // when a tile's raw maximum exceeds it by more than 2^8, so the accumulator in TMEM is rescaled a few times per
      // item instead of once per tile (D / l is invariant).  With the bias added, probabilities stay below
      // 2^(8 + max|bias|) -- far inside bf16 / fp32 range (fusion.py rejects bias tables beyond +-40).
      float m_c = -INFINITY, l_c = 0.f, m_o = -INFINITY, l_o = 0.f;
      for (int t = 0; t < ntiles; ++t, ++tcnt) {
        const int nval = min(kS, nv - t * kS);
        const uint32_t ko_u = smem_u32(sKoff + (tcnt & 3u) * kS);                      // -4 * koff of every key of the tile
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          const uint32_t tS = tm + lane_base + pr * 128, tD = tS + 64 + hh * 32;
          const uint32_t bt_u = smem_u32(sBias + (pr * 2 + hh) * kBiasStride + bias_q);
          if (tid == 0) FA_TS(0, ts_i++, 1);                    // step start (waiting for S)
          mbar_wait(&s_full[pr], tcnt & 1u);
          tc_fence_after();
          if (tid == 0) FA_TS(0, ts_i++, 2);                    // S available
          float lsum = 0.f;
          if (!(HMVIT_FA_DBG & 2)) {
            // The S row is read out of TMEM ONCE (TMEM -> register bandwidth is as scarce as MUFU throughput at head
            // dimension 32: 4 bytes per logit) and stays in registers for both passes.
            uint32_t sv[64];
            {
              uint32_t (&s0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
              uint32_t (&s1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);
              tmem_ld32(tS, s0);
              tmem_ld32(tS + 32, s1);
              tmem_ld_wait();
            }
            // ---- pass 1: maximum of the raw logits ----
            float mx = -INFINITY, mh = -INFINITY;
            if (nval < kS) {                                       // tail of the item's last tile
#pragma unroll
              for (int e = 0; e < 64; ++e) if (e >= nval) sv[e] = 0xff800000u;   // -inf
            }
#pragma unroll
            for (int e = 0; e < 64; e += 4) {
              mx = fmaxf(mx, fmaxf(__uint_as_float(sv[e]), __uint_as_float(sv[e + 1])));
              mh = fmaxf(mh, fmaxf(__uint_as_float(sv[e + 2]), __uint_as_float(sv[e + 3])));
            }
            mx = fmaxf(mx, mh);
            float alpha = 1.0f;
            if (mx > m_c + 8.0f) { alpha = ex2(m_c - mx); m_c = mx; }   // always on the first tile (m = -inf -> alpha = 0)
            l_c *= alpha;
            if (t > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) {
              // the reference moved: rescale this head's accumulator in TMEM, 16 columns at a time (P V of the previous
              // tile has retired: the commit behind s_full covers every earlier MMA)
#pragma unroll 1
              for (int h = 0; h < 2; ++h) {
                uint32_t d[16];
                tmem_ld16(tD + h * 16, d);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 16; ++k) d[k] = __float_as_uint(__uint_as_float(d[k]) * alpha);
                tmem_st16(tD + h * 16, d);
              }
            }
            if (tid == 0) FA_TS(0, ts_i++, 3);                  // pass 1 (+ rescale) done
            // ---- pass 2: + relative position bias, exp2, row sum, bf16 pack; four batches of 16 logits.  The two-level
            // bias lookup (key offset -> table word) must be ISSUED as batches -- left alone, ptxas sinks every load
            // next to its consumer (register pressure: 64 live logits) and the single warp then pays the full
            // shared-memory latency per logit (measured: IPC 0.12, short scoreboard on every FADD).  The warp-level
            // barriers between the phases are scheduling fences for the loads: 4 offset quads | 16 table words | math.
            // P is written back over the consumed logits (tail keys: the exponential of -inf is an exact 0). ----
            const float nmu = -m_c;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t kq[16], pk[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 x = lds_u4_addr(ko_u + c * 64 + q * 16);
                kq[q * 4 + 0] = x.x; kq[q * 4 + 1] = x.y; kq[q * 4 + 2] = x.z; kq[q * 4 + 3] = x.w;
              }
              __syncwarp();
#pragma unroll
              for (int e = 0; e < 16; ++e) kq[e] = __float_as_uint(lds_f32(bt_u + kq[e]));
              __syncwarp();
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float p0 = ex2(__uint_as_float(sv[c * 16 + 2 * k]) + (__uint_as_float(kq[2 * k]) + nmu));
                const float p1 = ex2(__uint_as_float(sv[c * 16 + 2 * k + 1]) + (__uint_as_float(kq[2 * k + 1]) + nmu));
                lsum += p0 + p1;
                pk[k] = pack_bf16x2(p0, p1);
              }
              tmem_st8(tS + c * 8, pk);
            }
          } else {
            uint32_t sa[32];
            tmem_ld32(tS, sa);
            tmem_ld_wait();
            tmem_st32(tS, sa);
          }
          l_c += lsum;
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[pr]);
          if (tid == 0) FA_TS(0, ts_i++, 9);                    // P handed over
          { const float tm_ = m_c, tl_ = l_c; m_c = m_o; l_c = l_o; m_o = tm_; l_o = tl_; }
        }
      }
      // ------------------------------ normalise and store ------------------------------
#pragma unroll 1
      for (int pr = 0; pr < 2; ++pr) {
        const uint32_t tD = tm + lane_base + pr * 128 + 64 + hh * 32;
        const int head = hgc * kHG + pr * 2 + hh;
        uint32_t d[32];
        float il = 0.f;
        if (ntiles > 0) {
          mbar_wait(&d_full[pr], icnt & 1u);
          tc_fence_after();
          if (tid == 0) FA_TS(0, ts_i++, 11);                      // accumulator final
          tmem_ld32(tD, d);
          tmem_ld_wait();
          if (tid == 0) FA_TS(0, ts_i++, 12);                      // accumulator in registers
          tc_fence_before();
          mbar_arrive(&d_free[pr]);                              // accumulator in registers: the next item may overwrite it
          il = l_c > 0.f ? 1.0f / l_c : 0.f;
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
        // training: softmax statistics (log2 domain: reference maximum + log2 of the denominator)
        if (p.lse != nullptr) p.lse[tok * kHeads + head] = l_c > 0.f ? m_c + log2f(l_c) : INFINITY;
        { const float tm_ = m_c, tl_ = l_c; m_c = m_o; l_c = l_o; m_o = tm_; l_o = tl_; }
      }
      if (ntiles > 0) ++icnt;
      if (tid == 0) FA_TS(0, ts_i++, 10);                           // item stored
    }
  } else if (warp < 12) {
    // =========================================== GATHER ===========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HMVIT_FA_GATHER_REGS));
    const int hw = (warp - 4) * 2 + (lane >> 4);        // half-warp 0..15: keys hw, hw + 16, hw + 32, hw + 48 of a tile
    const int u16 = lane & 15, pr = u16 >> 3, un = u16 & 7;
    const size_t plane = static_cast<size_t>(p.B) * p.L * N * 32;            // uint4 units per ego-type plane
    const uint32_t kv_u = smem_u32(sKV), kvb_u = smem_u32(sKvb);
    uint32_t tcnt = 0;
    int ts_i = 0; (void)ts_i;
    // Records and item sizes are requested one tile / one item ahead: lane q of a half-warp fetches the record of the
    // half-warp's q-th key of the NEXT tile (the slot exists for every item; past its visible keys the content is
    // stale and never used), so that neither latency sits on the tile's critical path.
    auto item_of = [&](int it_, int& a_, int& grp_) { a_ = sValid[it_ / G]; grp_ = it_ - (it_ / G) * G; };
    auto rec_ptr = [&](int a_, int grp_) {
      return reinterpret_cast<const uint4*>(fp.rec + (static_cast<size_t>(a_) * G + grp_) * (static_cast<size_t>(p.L) * kS));
    };
    const int mykey = hw + (u16 & 3) * 16;
    const int gt = tid - 128;                              // 0..255 among the gather threads
    // the item's 64 query rows (256 B of this head group each) pulled into L2 one item ahead
    auto prefetch_q = [&](int a_, int grp_) {
      if (gt < kS) {
        int r, c; group_token(p.kind, grp_ / GX, grp_ - (grp_ / GX) * GX, gt, p.H, p.W, r, c);
        const __nv_bfloat16* row = p.q + (static_cast<size_t>(a_) * N + r * p.W + c) * kC + hgc * 128;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(row) : "memory");
      }
    };
    uint4 rec_next = make_uint4(0, 0, 0, 0);
    uint32_t icnt = 0;
    if (item0 < n_items) {
      int a_, g_; item_of(item0, a_, g_);
      rec_next = __ldg(rec_ptr(a_, g_) + mykey);
      prefetch_q(a_, g_);
    }
    int kit = 0;
    for (int it = item0; it < n_items; it += item_step, ++kit) {
      int a, grp; item_of(it, a, grp);
      const int b = a / p.L;
      const int nv = item_nv(kit, it);
      const int ntiles = (nv + kS - 1) >> 6;
      const int it_n = it + item_step;
      int a_n = a, grp_n = grp;
      if (it_n < n_items) {
        item_of(it_n, a_n, grp_n);
        prefetch_q(a_n, grp_n);
      }
      if (ntiles == 0 && it_n < n_items) rec_next = __ldg(rec_ptr(a_n, grp_n) + mykey);
      const int te = p.mode[a] != 0 ? 1 : 0;
      const uint4* kbase = reinterpret_cast<const uint4*>(p.k) + te * plane + static_cast<size_t>(b) * p.L * N * 32 + hgc * 16 + u16;
      const uint4* vbase = reinterpret_cast<const uint4*>(p.v) + te * plane + static_cast<size_t>(b) * p.L * N * 32 + hgc * 16 + u16;
      const uint4* recs = rec_ptr(a, grp);
      const uint32_t kvb_te = kvb_u + te * 1024 + u16 * 16;
      for (int t = 0; t < ntiles; ++t, ++tcnt) {
        const int nval = min(kS, nv - t * kS);
        const uint32_t stage = tcnt & 1u;
        // this half-warp's 4 records of the tile (lane q holds record q), and the request for the next tile's
        const uint4 myrec = rec_next;
        if (t + 1 < ntiles) rec_next = __ldg(recs + (t + 1) * kS + mykey);
        else if (it_n < n_items) rec_next = __ldg(rec_ptr(a_n, grp_n) + mykey);
        if (tid == 128) FA_TS(1, ts_i++, 1);                        // tile start (waiting for the ring stage)
        mbar_wait(&kv_empty[stage], ((tcnt >> 1) & 1u) ^ 1u);      // P V of the tile two back has retired
        if (tid == 128) FA_TS(1, ts_i++, 2);                        // stage free
        const uint32_t dstK = kv_u + stage * Cfg::KV_STAGE + pr * 8192, dstV = dstK + 16384;
        int* koff = sKoff + (tcnt & 3u) * kS;

        // Everything on the tap path is volatile asm, i.e. issued in program order: left to themselves NVVM / ptxas sink
        // the tap loads next to their blend (56 registers), which serialises 32 L2 round trips per tile (measured: 8-10 k
        // cycles per tile instead of ~2 k).  Software pipeline over the half-warp's 4 keys: the value taps of key n are in
        // flight while its key taps are blended, the key taps of key n + 1 while its value taps are blended.
        uint4 kk[4], vv[4];
        uint32_t w01 = 0, w23 = 0, meta = 0;
        auto fetch = [&](int n) {                                  // record n of this half-warp -> tap weights, row offset
          const int src = (lane & 16) | n;
          const uint32_t r0 = __shfl_sync(0xffffffffu, myrec.x, src), r1 = __shfl_sync(0xffffffffu, myrec.y, src);
          const uint32_t r2 = __shfl_sync(0xffffffffu, myrec.z, src), r3 = __shfl_sync(0xffffffffu, myrec.w, src);
          const bool live = hw + n * 16 < nval;                    // past the item's visible keys: zero row, no loads
          w01 = live ? r1 : 0u; w23 = live ? r2 : 0u;
          meta = r3;
          return static_cast<int>(r0);                             // uint4 offset of the footprint corner's row in the scene's planes
        };
        auto issue = [&](const uint4* base, int off, uint32_t wa, uint32_t wb, uint4 (&tv)[4]) {
          // a tap outside the map has weight 0 and is not loaded
          tv[0] = ldg_nc_u4_if(base + off, (wa & 0xffffu) != 0u && !(HMVIT_FA_DBG & 1));
          tv[1] = ldg_nc_u4_if(base + off + 32, (wa >> 16) != 0u && !(HMVIT_FA_DBG & 1));
          tv[2] = ldg_nc_u4_if(base + off + p.W * 32, (wb & 0xffffu) != 0u && !(HMVIT_FA_DBG & 1));
          tv[3] = ldg_nc_u4_if(base + off + p.W * 32 + 32, (wb >> 16) != 0u && !(HMVIT_FA_DBG & 1));
        };
        auto blend_store = [&](const uint4 (&tv)[4], uint32_t wa, uint32_t wb, uint32_t bias_addr, uint32_t dst, int key) {
          uint4 o = lds_u4_addr(bias_addr);
          const uint32_t w2[4] = {__byte_perm(wa, wa, 0x1010), __byte_perm(wa, wa, 0x3232), __byte_perm(wb, wb, 0x1010),
                                  __byte_perm(wb, wb, 0x3232)};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            o.x = hfma2_bf16_v(w2[q], tv[q].x, o.x); o.y = hfma2_bf16_v(w2[q], tv[q].y, o.y);
            o.z = hfma2_bf16_v(w2[q], tv[q].z, o.z); o.w = hfma2_bf16_v(w2[q], tv[q].w, o.w);
          }
          if (key >= nval) o = make_uint4(0, 0, 0, 0);             // (selects, no branch)
          sts_u4_addr(dst + sw128_offset(key, un), o);
        };
        int off = fetch(0);
        issue(kbase, off, w01, w23, kk);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const int key = hw + n * 16;
          const uint32_t tj = (meta >> 16) & 1u, ko = (meta >> 24) * 4u;
          const uint32_t wa = w01, wb = w23;
          issue(vbase, off, wa, wb, vv);
          blend_store(kk, wa, wb, kvb_te + tj * 512, dstK, key);
          if (u16 == 0) koff[key] = key < nval ? -static_cast<int>(ko) : 0;
          if (n + 1 < 4) { off = fetch(n + 1); issue(kbase, off, w01, w23, kk); }
          blend_store(vv, wa, wb, kvb_te + tj * 512 + 256, dstV, key);    // the weights of key n (w01 / w23 now hold key n + 1's)
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_full[stage]);
        if (tid == 128) FA_TS(1, ts_i++, 3);                        // tile delivered
        if (t == 0) {
          // the item's 64 query rows -> block-diagonal Q tiles (64 tokens x 16 units = 4 units per gather thread), once
          // the previous item's last Q K^T has retired; behind tile 0 so that the MMA warp finds both when it arrives
          const int gy = grp / GX, gx = grp - gy * GX;
          const uint4* qsrc = reinterpret_cast<const uint4*>(p.q) + static_cast<size_t>(a) * N * 32 + hgc * 16;
          uint4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int idx = gt + k * 256;
            int r, c; group_token(p.kind, gy, gx, idx >> 4, p.H, p.W, r, c);
            v[k] = __ldg(qsrc + static_cast<size_t>(r * p.W + c) * 32 + (idx & 15));
          }
          mbar_wait(q_empty, (icnt & 1u) ^ 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int idx = gt + k * 256;
            const int sq = idx >> 4, u = idx & 15;
            const int prq = u >> 3, hq = (u >> 2) & 1, cq = u & 3;
            *reinterpret_cast<uint4*>(sQ + prq * 16384 + sw128_offset(hq * 64 + sq, hq * 4 + cq)) = v[k];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(q_full);
          if (tid == 128) FA_TS(1, ts_i++, 4);                      // Q delivered
          ++icnt;
        }
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
    int ts_i = 0; (void)ts_i;
    int kit = 0;
    for (int it = item0; it < n_items; it += item_step, ++kit) {
      const int nv = item_nv(kit, it);
      const int ntiles = (nv + kS - 1) >> 6;
      if (ntiles == 0) continue;
      if (lane == 0) FA_TS(2, ts_i++, 1);                           // item start (waiting for tile 0)
      mbar_wait(&kv_full[tcnt & 1u], (tcnt >> 1) & 1u);
      if (lane == 0) FA_TS(2, ts_i++, 2);                           // tile 0 there (waiting for Q)
      mbar_wait(q_full, icnt & 1u);
      if (lane == 0) FA_TS(2, ts_i++, 3);                           // Q there
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
        const uint32_t nph = ((tcnt + 1) >> 1) & 1u;
        // The next tile's Q K^T of pair 0 is issued right behind this tile's P V of pair 0 (the softmax warpgroup is then
        // busy with pair 1) -- but only if the gather warps have already delivered that tile: this tile's P V and the
        // release of its ring stage must not wait for them.
        bool qk0_done = false;
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          mbar_wait(&p_full[pr], tcnt & 1u);
          if (lane == 0) FA_TS(2, ts_i++, 4);                       // P of this pair there
          if (t == 0) mbar_wait(&d_free[pr], (icnt & 1u) ^ 1u);   // the previous item's accumulator has been read
          tc_fence_after();
          // (lane 0's view, so that the whole warp takes the same path)
          const bool next_ready = has_next && (pr == 1 ? qk0_done
                                  : __shfl_sync(0xffffffffu, mbar_test_wait(&kv_full[stage ^ 1u], nph) ? 1 : 0, 0) != 0);
          if (next_ready) tc_fence_after();
          if (elect_one()) {
            issue_pv(pr, stage, t == 0);
            if (!has_next) umma_commit(&d_full[pr]);
            if (pr == 1) umma_commit(&kv_empty[stage]);           // both pairs' P V of this tile issued
            if (next_ready) { issue_qk(pr, stage ^ 1u); umma_commit(&s_full[pr]); }
          }
          __syncwarp();
          if (pr == 0) qk0_done = next_ready;
        }
        if (has_next) {
          if (!qk0_done) {
            mbar_wait(&kv_full[stage ^ 1u], nph);
            tc_fence_after();
          }
          if (elect_one()) {
            if (!qk0_done) {
              issue_qk(0, stage ^ 1u); umma_commit(&s_full[0]);
              issue_qk(1, stage ^ 1u); umma_commit(&s_full[1]);
            }
            if (t + 2 == ntiles) umma_commit(q_empty);            // the item's last Q K^T has been issued
          }
          __syncwarp();
        }
      }
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
