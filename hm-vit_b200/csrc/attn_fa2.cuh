// Fused warp + mask + multi-agent group attention, second form of the persistent tcgen05 kernel (default of
// hmvit_group_attn / hmvit_fusion_forward).  Same contract and the same key records as attn_fused.cuh; replaces
//   HeteroFusionBlock.warp_features            hetero_fusion.py:338-361
//   the ego loop around HeteroAttention.forward hetero_fusion.py:373-397 / 412-440, 187-277
//   relative position bias                      hetero_fusion.py:82-109, 227-237
//
// What changed against the first form (profiles/r2_ncu_full_mid.md: L1 data pipe 73 % busy, ~10 instructions per logit,
// one softmax warp per SM sub-partition and CTA, 0.33 ms of pure skeleton per launch):
//   * ONE persistent CTA per SM (896 threads x 72 registers), still one head group (4 heads = 2 head PAIRS) per CTA.
//   * The relative position bias is part of the Q K^T contraction: the key tile carries a 64-wide ONE-HOT column block
//     (fp16, which group slot the compacted key came from) and the A operand a static 128 x 64 fp16 bias matrix
//     (row (head, query) x slot), so  S = Qbd K^T + Bias OneHot^T  leaves the tensor core with the bias already added.
//     The softmax threads no longer look anything up (the old form spent 4 of its 10 instructions per logit and 30 % of
//     the L1 data pipe on the two-level bias lookup).  fp16 bias: 11-bit significand, |error| <= 2^-12 |bias|.
//   * The block-diagonal Q lives in TENSOR MEMORY (written by the softmax threads with tcgen05.st), so the Q K^T A operand
//     never crosses the shared-memory port.
//   * S is DOUBLE BUFFERED per head pair in tensor memory: Q K^T of tile t + 1 is computed while the pair's softmax
//     warpgroup works on tile t (P overwrites S, so with one buffer the chain  S -> softmax -> P -> P V -> next S  is
//     serial and each warpgroup idles ~40 % of the time).  One MMA warp per pair, each a small state machine over two
//     cursors (Q K^T up to two tiles ahead of P V) instead of a fixed issue order.
//   * Two softmax warpgroups, one per head pair, run concurrently (thread == TMEM lane == (head of the pair, query row)).
//     The S row is read from tensor memory twice, 32 columns at a time (maximum pass, exponential pass): 72 registers.
//   * 16 gather warps (half-warp == key, 2 keys per tile).  They are bound by their own dependent instruction stream
//     (~4 cycles per instruction and warp), so what counts is the warp count and a short per-key path: the record pass
//     stores ready-made row offsets with the footprint corner moved into the map (all four taps load unconditionally),
//     a loader warp bulk-copies each tile's 64 records into a shared-memory ring (one broadcast LDS.128 per key), the
//     per-thread address terms are hoisted.
//   * The last tile of an item is issued with N = visible keys rounded up to 16 and the softmax skips dead chunks.  The
//     up to 15 padding rows of that tile REPEAT the item's last visible key (same key row, same one-hot row) with a zero
//     value row, so neither the maximum nor P V needs a mask; the softmax denominator comes out of the tensor core too
//     (l = P live^T with a K-major [16 x keys] indicator tile whose row 0 is 1 for visible keys), i.e. no per-logit add
//     and no tail select in the softmax threads: ~3 instructions per logit (max, subtract, exp2, half a pack).
#pragma once
#include "attn_fused.cuh"

namespace hmvit {

// registers per thread after the role split (896 threads launch with 72 each = 64512): spill reloads sit on the roles'
// critical paths and miss the small L1 (local memory competes with the tap stream), so the softmax warps get what their
// tile loop needs, the gather warps exactly what theirs needs
#ifndef HMVIT_FA2_SOFT_REGS
#define HMVIT_FA2_SOFT_REGS 88
#endif
#ifndef HMVIT_FA2_GATHER_REGS
#define HMVIT_FA2_GATHER_REGS 72
#endif
#ifndef HMVIT_FA2_MISC_REGS
#define HMVIT_FA2_MISC_REGS 40
#endif
#ifndef HMVIT_FA2_STAGES
#define HMVIT_FA2_STAGES 3
#endif
#ifndef HMVIT_FA2_AHEAD      // tiles Q K^T may run ahead of P V per pair (2 = both S buffers; 1 = serial, debugging)
#define HMVIT_FA2_AHEAD 2
#endif

struct Fa2Cfg {
  static constexpr int SOFT_WARPS = 8;                  // warpgroup 0: pair 0, warpgroup 1: pair 1
  static constexpr int GATHER_WARPS = 16;
  static constexpr int MMA_WARP = SOFT_WARPS + GATHER_WARPS;
  static constexpr int THREADS = (SOFT_WARPS + GATHER_WARPS + 4) * 32;      // 896; the last warpgroup: MMA warp of pair 0, loader warp, MMA warp of pair 1, 1 idle
  static constexpr int STAGES = HMVIT_FA2_STAGES;       // ring depth
  static constexpr int STAGE_BYTES = 43008;             // K: 2 pairs x 8 KB | V: 2 pairs x 8 KB | one-hot: 8 KB | live indicator: 2 KB
  static constexpr int OFF_RING = 0;
  static constexpr int OFF_AUG = STAGES * STAGE_BYTES;                   // [2 pairs][128 rows][64 slots] fp16 bias matrix (SW128 K-major A operand)
  static constexpr int OFF_KVB = OFF_AUG + 2 * 16384;                    // [2 te][2 tj][K | V][128 ch] bf16 folded biases
  static constexpr int OFF_BIAS = OFF_KVB + 2 * 2 * 2 * 256;             // [4 heads][kBiasStride] fp32, log2 domain (prologue)
  static constexpr int OFF_VALID = OFF_BIAS + kHG * kBiasStride * 4;     // [kFusedMaxAgents] uint16
  static constexpr int OFF_NV = OFF_VALID + kFusedMaxAgents * 2;         // [NV_ITEMS] uint16
  static constexpr int NV_ITEMS = 2048;
  static constexpr int OFF_ITEM = OFF_NV + NV_ITEMS * 2;                  // [NV_ITEMS] uint32: agent << 16 | group of this CTA's k-th item
  static constexpr int OFF_OH = OFF_ITEM + NV_ITEMS * 4;                  // [64 slots][64] fp16 one-hot rows (static)
  static constexpr int REC_STAGES = 8;                                    // ring of the tiles' key records (1 KB each)
  static constexpr int OFF_REC = OFF_OH + 64 * 128;
  static constexpr int OFF_BAR = OFF_REC + REC_STAGES * 1024;
  static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;                // + alignment slack
  // tensor memory columns
  static constexpr uint32_t TM_COLS = 512;
  static constexpr uint32_t TM_S = 0;                   // pair p, buffer b: S / P at 128 p + 64 b
  static constexpr uint32_t TM_D = 256;                 // pair p: D at 256 + 64 p (head hh of the pair: + 32 hh)
  static constexpr uint32_t TM_Q = 384;                 // pair p: block-diagonal Q (bf16, 32 columns) at 384 + 32 p
  static constexpr uint32_t TM_L = 448;                 // pair p: softmax denominators (column 0 of 16) at 448 + 16 p
};

// position in a CTA's sequence of key tiles (its non-empty items, in order)
struct Fa2Cur { int kit, it, t, nv; };

HMVIT_DEVINL float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__global__ void __launch_bounds__(Fa2Cfg::THREADS, 1) fused_attn2_kernel(const FusedAttnParams fp) {
  using Cfg = Fa2Cfg;
  const AttnParams& p = fp.a;
  const int N = p.H * p.W;
  const int GX = p.W / kWin, G = (p.H / kWin) * GX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sRing = smem + Cfg::OFF_RING;
  uint8_t* sKvb = smem + Cfg::OFF_KVB;
  float* sBias = reinterpret_cast<float*>(smem + Cfg::OFF_BIAS);
  uint16_t* sValid = reinterpret_cast<uint16_t*>(smem + Cfg::OFF_VALID);
  uint16_t* sNv = reinterpret_cast<uint16_t*>(smem + Cfg::OFF_NV);
  uint32_t* sItem = reinterpret_cast<uint32_t*>(smem + Cfg::OFF_ITEM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* kv_full = bars + 0;      // [<= 4 stages] gather warps -> MMA
  uint64_t* kv_empty = bars + 4;     // [<= 4 stages] MMA (commit) -> gather warps
  uint64_t* s_full = bars + 8;       // [2 pairs][2 buffers]  MMA (commit) -> softmax
  uint64_t* p_full = bars + 12;      // [2 pairs][2 buffers]  softmax -> MMA
  uint64_t* d_full = bars + 16;      // [2 pairs]   MMA (commit) -> softmax: the item's accumulator is final
  uint64_t* d_free = bars + 18;      // [2 pairs]   softmax -> MMA: accumulator read, the next item may overwrite it
  uint64_t* q_full = bars + 20;      // [2 pairs]   softmax -> MMA: the pair's block-diagonal Q of the next item is in TMEM
  uint64_t* rec_full = bars + 22;    // [8 stages]  loader warp (bulk copy) -> gather warps: the tile's key records in shared memory
  uint64_t* rec_empty = bars + 30;   // [8 stages]  gather warps -> loader warp: records read into registers
  uint64_t* pv_done = bars + 38;     // [2 pairs]   MMA (commit behind every P V) -> softmax: the accumulator may be rescaled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
  int* s_nvalid = reinterpret_cast<int*>(bars + 41);
  static_assert(Cfg::STAGES >= 2 && Cfg::STAGES <= 4, "ring depth");
  static_assert(Cfg::SMEM_BYTES <= 232448, "shared memory");

  // ------------------------------ one-off prologue ------------------------------
  const int hgc = blockIdx.x & 1;                       // head group of this CTA
  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&kv_full[s], Cfg::GATHER_WARPS); mbar_init(&kv_empty[s], 2); }
    for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&d_full[s], 1); mbar_init(&d_free[s], 128); mbar_init(&q_full[s], 128); mbar_init(&pv_done[s], 1);
    }
    for (int s = 0; s < Cfg::REC_STAGES; ++s) { mbar_init(&rec_full[s], 1); mbar_init(&rec_empty[s], Cfg::GATHER_WARPS); }
    fence_mbar_init();
  }
  if (warp == Cfg::MMA_WARP) tmem_alloc<Cfg::TM_COLS>(tmem_slot);
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
  for (int e = tid; e < 64 * 32; e += Cfg::THREADS)        // one-hot rows: 1.0 (fp16) at the slot's position
    reinterpret_cast<uint32_t*>(smem + Cfg::OFF_OH)[e] = (e & 31) == ((e >> 5) >> 1) ? (((e >> 5) & 1) ? 0x3c000000u : 0x00003c00u) : 0u;
  for (int e = tid; e < Cfg::STAGES * 128; e += Cfg::THREADS)   // live-indicator tiles: rows 1..15 stay zero, row 0 is written per tile
    *reinterpret_cast<uint4*>(sRing + (e >> 7) * Cfg::STAGE_BYTES + 40960 + (e & 127) * 16) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  // bias matrix A operand: row (pair pr, head hh, query row) x slot r -> table[bias_q - koff(r)], koff = (r >> 3) * 15 + (r & 7)
  if (tid < 256) {
    const int pr = tid >> 7, hh = (tid >> 6) & 1, row = tid & 63;
    const float* bt = sBias + (pr * 2 + hh) * kBiasStride + ((row >> 3) + 7) * 15 + (row & 7) + 7;
    uint8_t* dstp = smem + Cfg::OFF_AUG + pr * 16384;
#pragma unroll 1
    for (int u = 0; u < 8; ++u) {                        // unit u = slots 8 u .. 8 u + 7 = window row u of the key
      uint4 w;
      w.x = pack_f16x2(bt[-u * 15 - 0], bt[-u * 15 - 1]); w.y = pack_f16x2(bt[-u * 15 - 2], bt[-u * 15 - 3]);
      w.z = pack_f16x2(bt[-u * 15 - 4], bt[-u * 15 - 5]); w.w = pack_f16x2(bt[-u * 15 - 6], bt[-u * 15 - 7]);
      *reinterpret_cast<uint4*>(dstp + sw128_offset(hh * 64 + row, u)) = w;
    }
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
  // visible-key counts of this CTA's items, staged once (a per-item global load would sit on every role's critical path)
  for (int k = tid; k < Cfg::NV_ITEMS; k += Cfg::THREADS) {
    const int it_ = item0 + k * item_step;
    if (it_ >= n_items) break;
    const int a_ = sValid[it_ / G], g_ = it_ - (it_ / G) * G;
    sItem[k] = (static_cast<uint32_t>(a_) << 16) | static_cast<uint32_t>(g_);
    sNv[k] = static_cast<uint16_t>(__ldg(fp.nvis + static_cast<size_t>(a_) * G + g_));
  }
  __syncthreads();
  auto item_nv = [&](int k_, int it_) -> int {                    // visible keys of this CTA's k-th item (= item it_)
    if (k_ < Cfg::NV_ITEMS) return sNv[k_];
    return __ldg(fp.nvis + static_cast<size_t>(sValid[it_ / G]) * G + (it_ - (it_ / G) * G));
  };
  // (agent, group) of this CTA's k-th item: from the table (the divisions by G were 1-2 k cycles per item on every role)
  auto item_ag = [&](int k_, int it_, int& a_, int& g_) {
    if (k_ < Cfg::NV_ITEMS) { const uint32_t v = sItem[k_]; a_ = static_cast<int>(v >> 16); g_ = static_cast<int>(v & 0xffffu); }
    else { a_ = sValid[it_ / G]; g_ = it_ - (it_ / G) * G; }
  };
  // first non-empty item at or after (kit_, it_)
  auto seek = [&](int& kit_, int& it_, int& nv_) {
    while (it_ < n_items) {
      nv_ = item_nv(kit_, it_);
      if (nv_ > 0) return;
      it_ += item_step; ++kit_;
    }
  };

  if (warp < Cfg::SOFT_WARPS) {
    // =========================================== SOFTMAX ===========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HMVIT_FA2_SOFT_REGS));
    static_assert(256 * HMVIT_FA2_SOFT_REGS + 512 * HMVIT_FA2_GATHER_REGS + 128 * HMVIT_FA2_MISC_REGS <= 896 * 72, "register pool");
    static_assert(HMVIT_FA2_SOFT_REGS >= 72 && HMVIT_FA2_GATHER_REGS <= 72 && HMVIT_FA2_MISC_REGS <= 72, "setmaxnreg direction");
    const int pr = warp >> 2;                             // head pair of this warpgroup
    const int hh = (tid >> 6) & 1, row = tid & 63;        // head of the pair, query row
    const int head = hgc * kHG + pr * 2 + hh;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS0 = tm + lane_base + Cfg::TM_S + pr * 128;   // + 64 buffer
    const uint32_t tD = tm + lane_base + Cfg::TM_D + pr * 64 + hh * 32;
    const uint32_t tQ = tm + lane_base + Cfg::TM_Q + pr * 32;
    const uint32_t tL = tm + lane_base + Cfg::TM_L + pr * 16;
    int ts_i = 0; (void)ts_i;
    {
      // static zero half of the block-diagonal Q: head 0 of the pair uses K-columns [0,32) = TMEM columns [0,16),
      // head 1 K-columns [32,64) = TMEM columns [16,32)
      uint32_t z[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) z[k] = 0u;
      tmem_st16(tQ + (1 - hh) * 16, z);
    }
    // this thread's query row of item it_: 64 B (its head's 32 channels)
    auto load_q = [&](int k_, int it_, uint4 (&qv)[4]) {
      int a_, g_; item_ag(k_, it_, a_, g_);
      int r, c; group_token(p.kind, g_ / GX, g_ - (g_ / GX) * GX, row, p.H, p.W, r, c);
      const uint4* src = reinterpret_cast<const uint4*>(p.q + (static_cast<size_t>(a_) * N + r * p.W + c) * kC + head * kDh);
#pragma unroll
      for (int k = 0; k < 4; ++k) qv[k] = __ldg(src + k);
    };
    auto store_q = [&](const uint4 (&qv)[4]) {
      uint32_t w[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) { w[4 * k] = qv[k].x; w[4 * k + 1] = qv[k].y; w[4 * k + 2] = qv[k].z; w[4 * k + 3] = qv[k].w; }
      tmem_st16(tQ + hh * 16, w);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&q_full[pr]);
    };
    {
      int k0 = 0, i0 = item0, n0 = 0;
      seek(k0, i0, n0);
      if (i0 < n_items) { uint4 qv[4]; load_q(k0, i0, qv); store_q(qv); }   // (also publishes the zero half)
    }
    uint32_t g = 0, ord = 0;                              // this pair's tile counter, non-empty item ordinal
    int kit = 0;
    for (int it = item0; it < n_items; it += item_step, ++kit) {
      int a, grp; item_ag(kit, it, a, grp);
      const int nv = item_nv(kit, it);
      const int ntiles = (nv + kS - 1) >> 6;
      int r, c; group_token(p.kind, grp / GX, grp - (grp / GX) * GX, row, p.H, p.W, r, c);
      const size_t tok = static_cast<size_t>(a) * N + r * p.W + c;
      // m: the maximum the exponentials refer to.  It follows the running maximum lazily: it only moves when a tile's
      // maximum exceeds it by more than 2^8, so the accumulator in TMEM is rescaled a few times per item at most
      // (D / l is invariant); probabilities stay below 2^8.
      float m_c = -INFINITY;
      for (int t = 0; t < ntiles; ++t, ++g) {
        const int nval = min(kS, nv - t * kS);
        const int nch = (nval + 15) >> 4;                                         // live 16-logit chunks
        const uint32_t buf = g & 1u;
        const uint32_t tS = tS0 + buf * 64;
        if (tid == 0) FA_TS(0, ts_i++, 1);                      // step start (waiting for S)
        bool waited = false;
        if (t == ntiles - 1) {
          // Last tile of the item: once its Q K^T has retired (the commit behind s_full covers every earlier MMA) the
          // pair's Q columns are free for the NEXT non-empty item's query row, requested before the wait.
          int kn = kit + 1, in_ = it + item_step, nn = 0;
          seek(kn, in_, nn);
          if (in_ < n_items) {
            uint4 qv[4];
            load_q(kn, in_, qv);
            mbar_wait_sleepy(&s_full[pr * 2 + buf], (g >> 1) & 1u);
            store_q(qv);
            waited = true;
          }
        }
        if (!waited) mbar_wait_sleepy(&s_full[pr * 2 + buf], (g >> 1) & 1u);
        tc_fence_after();
        if (tid == 0) FA_TS(0, ts_i++, 2);                      // S available
        // The S row is read from tensor memory TWICE, 32 columns at a time (maximum pass, exponential pass): 32 live
        // logits instead of 64 keep the softmax threads at 72 registers, which is what lets 16 gather warps run beside them.
        uint32_t sv[32];
        // ---- maximum of the logits (padding rows repeat a visible key: no mask) ----
        float mx = -INFINITY;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf * 2 < nch && !(HMVIT_FA_DBG & 2)) {
            tmem_ld32(tS + hf * 32, sv);
            tmem_ld_wait();
#pragma unroll
            for (int cch = 0; cch < 2; ++cch) {
              if (hf * 2 + cch < nch) {
                float m0 = mx, m1 = -INFINITY;
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                  m0 = fmax3(m0, __uint_as_float(sv[cch * 16 + e]), __uint_as_float(sv[cch * 16 + e + 1]));
                  m1 = fmax3(m1, __uint_as_float(sv[cch * 16 + e + 2]), __uint_as_float(sv[cch * 16 + e + 3]));
                }
                mx = fmaxf(m0, m1);
              }
            }
          }
        }
        float alpha = 1.0f;
        if (mx > m_c + 8.0f) { alpha = ex2(m_c - mx); m_c = mx; }   // always on the first tile (m = -inf -> alpha = 0)
        if (t > 0 && !__all_sync(0xffffffffu, alpha == 1.0f)) {
          // The reference moved: rescale this row's accumulator and denominator in TMEM.  With two S buffers Q K^T of this
          // tile may have been issued BEFORE P V of the previous one, so s_full does not certify that the accumulator is
          // complete: pv_done does (committed behind every P V of the pair; phase == tile counter).  P V of THIS tile
          // cannot start before the P handed over below.
          mbar_wait(&pv_done[pr], (g - 1u) & 1u);
          tc_fence_after();
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t d[16];
            tmem_ld16(tD + h * 16, d);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) d[k] = __float_as_uint(__uint_as_float(d[k]) * alpha);
            tmem_st16(tD + h * 16, d);
          }
          uint32_t lv[1];
          tmem_ld1(tL, lv);
          tmem_ld_wait();
          lv[0] = __float_as_uint(__uint_as_float(lv[0]) * alpha);
          tmem_st1(tL, lv);
        }
        if (tid == 0) FA_TS(0, ts_i++, 3);                      // maximum (+ rescale) done
        // ---- exp2, bf16 pack; P (32 packed columns) is written back over the consumed logits: the first half's P lands on
        // S columns [0,16), the second half's on [16,32), both already in registers when they are overwritten ----
        const float nmu = -m_c;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          if (hf * 2 < nch && !(HMVIT_FA_DBG & 2)) {
            tmem_ld32(tS + hf * 32, sv);
            tmem_ld_wait();
#pragma unroll
            for (int cch = 0; cch < 2; ++cch) {
              if (hf * 2 + cch < nch) {
                uint32_t pk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const float p0 = ex2(__uint_as_float(sv[cch * 16 + 2 * k]) + nmu);
                  const float p1 = ex2(__uint_as_float(sv[cch * 16 + 2 * k + 1]) + nmu);
                  pk[k] = pack_bf16x2(p0, p1);
                }
                tmem_st8(tS + hf * 16 + cch * 8, pk);
              }
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&p_full[pr * 2 + buf]);
        if (tid == 0) FA_TS(0, ts_i++, 9);                      // P handed over
      }
      // ------------------------------ normalise and store ------------------------------
      {
        uint32_t d[32];
        float il = 0.f, l_c = 0.f;
        if (ntiles > 0) {
          mbar_wait_sleepy(&d_full[pr], ord & 1u);
          tc_fence_after();
          if (tid == 0) FA_TS(0, ts_i++, 11);                   // accumulator final
          uint32_t lv[1];
          tmem_ld32(tD, d);
          tmem_ld1(tL, lv);
          tmem_ld_wait();
          l_c = __uint_as_float(lv[0]);
          tc_fence_before();
          mbar_arrive(&d_free[pr]);                              // accumulator in registers: the next item may overwrite it
          il = l_c > 0.f ? 1.0f / l_c : 0.f;
          ++ord;
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
      }
      if (tid == 0) FA_TS(0, ts_i++, 10);                         // item stored
    }
  } else if (warp < Cfg::MMA_WARP) {
    // =========================================== GATHER ===========================================
#if HMVIT_FA2_GATHER_REGS < 72
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HMVIT_FA2_GATHER_REGS));
#endif
    // The tile loop must not spill (local memory is larger than what is left of L1 beside 200 KB of shared memory, so
    // every spill reload is an L2 round trip on the tile's critical path) and must be SHORT: the gather warps are bound
    // by their own dependent instruction stream (~4 cycles per instruction and warp), not by the memory system: 16 warps,
    // 72 registers each.  The
    // record pass stores ready-made row offsets, the loader warp puts a tile's 64 records into shared memory (one
    // broadcast LDS.128 per key instead of shuffles + unpacking), and all per-thread address terms are hoisted.
    const int hw = (warp - Cfg::SOFT_WARPS) * 2 + (lane >> 4);   // half-warp 0..31: keys hw, hw + 32 of a tile
    const int u16 = lane & 15;                                    // 16-byte unit of the head group's 256-byte row
    const uint32_t ring_u = smem_u32(sRing), kvb_u = smem_u32(sKvb) + u16 * 16, oh_u = smem_u32(smem + Cfg::OFF_OH) + u16 * 16;
    const uint32_t rec_u = smem_u32(smem + Cfg::OFF_REC);
    // swizzled position of this lane's unit in a key row: key = hw + 32 n has (key & 7) == (hw & 7)
    const uint32_t row_c = (u16 >> 3) * 8192u + static_cast<uint32_t>(hw) * 128u + (((u16 & 7) ^ (hw & 7)) << 4);
    const int gt = tid - Cfg::SOFT_WARPS * 32;             // 0..511 among the gather threads
    const size_t wrow = static_cast<size_t>(p.W) * 32;     // one map row in 16-byte units
    uint32_t tcnt = 0;
    int ts_i = 0; (void)ts_i;
    int kit = 0;
    for (int it = item0; it < n_items; it += item_step, ++kit) {
      const int nv = item_nv(kit, it);
      const int ntiles = (nv + kS - 1) >> 6;
      if (ntiles == 0) continue;
      const uint4* kbase; const uint4* vbase;
      uint32_t kvb_te;
      {
        int a, grp_; item_ag(kit, it, a, grp_);
        const int b = a / p.L;
        const int te = p.mode[a] != 0 ? 1 : 0;
        // (ego type plane, scene, this lane's unit) in the K' / V' planes
        const size_t base_off = (static_cast<size_t>(te) * p.B + b) * static_cast<size_t>(p.L) * N * 32u + hgc * 16 + u16;
        kbase = reinterpret_cast<const uint4*>(p.k) + base_off; vbase = reinterpret_cast<const uint4*>(p.v) + base_off;
        kvb_te = kvb_u + te * 1024;
      }
      for (int t = 0; t < ntiles; ++t, ++tcnt) {
        const int nval = min(kS, nv - t * kS);
        const int nk = (nval + 15) & ~15;                          // key rows the MMAs of this tile read
        const int nkeys = hw < nk ? (hw + 32 < nk ? 2 : 1) : 0;    // keys of this half-warp in the tile (uniform per warp: nk is a multiple of 16)
        const uint32_t stage = tcnt % Cfg::STAGES, rst = tcnt & (Cfg::REC_STAGES - 1);
        if (gt == 0) FA_TS(1, ts_i++, 1);                           // tile start (waiting for the records)
        // this half-warp's records of the tile (padding rows repeat the last visible key)
        uint4 rc[2];
        mbar_wait(&rec_full[rst], (tcnt / Cfg::REC_STAGES) & 1u);
        if (gt == 0) FA_TS(1, ts_i++, 4);                           // records there
#pragma unroll
        for (int n = 0; n < 2; ++n)
          if (n < nkeys) rc[n] = lds_u4_addr(rec_u + rst * 1024 + min(hw + n * 32, nval - 1) * 16);
        // Tap loads are volatile asm, i.e. issued in program order.  The half-warp's 2 keys are 4 tap GROUPS (key taps,
        // value taps: 4 rows of 256 B each); two groups are in flight at any time (32 registers; 32 half-warps x 2 KB =
        // 64 KB per SM).
        uint4 kk[4], vv[4];
        // A key whose footprint is exactly one pixel (weights 1, 0, 0, 0: the ego's own keys, a third of all keys, and
        // whole-cell poses) loads one row instead of four and adds it to the folded bias.
        auto is_single = [](const uint4& r) { return r.y == 0x00003f80u && r.z == 0u; };
        auto issue = [&](const uint4* base, const uint4& r, uint4 (&tv)[4]) {
          // all four taps are in the map (the record pass moved the corner; outside taps carry weight 0)
          const uint4* r0 = base + r.x;
          const uint4* r1 = r0 + wrow;
          if (HMVIT_FA_DBG & 1) { tv[0] = tv[1] = tv[2] = tv[3] = make_uint4(0, 0, 0, 0); return; }
          tv[0] = ldg_nc_u4_v(r0);
          if (!is_single(r)) {
            tv[1] = ldg_nc_u4_v(r0 + 32);
            tv[2] = ldg_nc_u4_v(r1);
            tv[3] = ldg_nc_u4_v(r1 + 32);
          }
        };
        auto blend_store = [&](const uint4 (&tv)[4], const uint4& r, uint32_t bias_addr, uint32_t daddr, bool live) {
          uint4 o = lds_u4_addr(bias_addr);
          if (is_single(r)) {
            const uint32_t one = 0x3f803f80u;
            o.x = hfma2_bf16(one, tv[0].x, o.x); o.y = hfma2_bf16(one, tv[0].y, o.y);
            o.z = hfma2_bf16(one, tv[0].z, o.z); o.w = hfma2_bf16(one, tv[0].w, o.w);
          } else {
            const uint32_t wa = r.y, wb = r.z;
            const uint32_t w2[4] = {__byte_perm(wa, wa, 0x1010), __byte_perm(wa, wa, 0x3232), __byte_perm(wb, wb, 0x1010),
                                    __byte_perm(wb, wb, 0x3232)};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              o.x = hfma2_bf16(w2[q], tv[q].x, o.x); o.y = hfma2_bf16(w2[q], tv[q].y, o.y);
              o.z = hfma2_bf16(w2[q], tv[q].z, o.z); o.w = hfma2_bf16(w2[q], tv[q].w, o.w);
            }
          }
          if (!live) o = make_uint4(0, 0, 0, 0);                   // (selects, no branch)
          sts_u4_addr(daddr, o);
        };
        // The ring stage is awaited BEFORE the first tap loads and before the records slot is released.  With the records
        // coming from the shared-memory ring, either of the two the other way round (tap loads issued while parked on
        // kv_empty; slot released -- and refilled by the loader -- while parked) produced rare wrong key rows in single
        // tiles, although no value written or read moves across the wait (a build that prefetched the records into
        // registers instead was exact with early loads but spilled: +30 %; a consumer-side proxy fence behind the records wait
        // did not help).  Not understood; the order that is
        // bit-reproducible over repeated runs is kept (tools/fa_stress.py).
        mbar_wait(&kv_empty[stage], ((tcnt / Cfg::STAGES) & 1u) ^ 1u);   // the tile STAGES back has been consumed
        if (nkeys > 0) { issue(kbase, rc[0], kk); issue(vbase, rc[0], vv); }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rec_empty[rst]);               // records in registers
        if (gt == 0) FA_TS(1, ts_i++, 2);                           // stage free
        const uint32_t sbase = ring_u + stage * Cfg::STAGE_BYTES;
        const uint32_t dst = sbase + row_c;                         // K row of key hw; + 4096 n: key hw + 32 n;
                                                                    // + 16384: V; (u16 < 8) + 32768: one-hot
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          if (n < nkeys) {                                         // (uniform)
            const uint32_t meta = rc[n].w;
            const uint32_t tjo = ((meta >> 16) & 1u) * 512u, slot = (meta >> 8) & 63u;
            const bool live = hw + n * 32 < nval;                  // padding row: key row repeated, value row zero
            blend_store(kk, rc[n], kvb_te + tjo, dst + n * 4096, true);
            if (u16 < 8) {
              // one-hot column block of the key (1.0 fp16 at its group slot): copied from the static table
              sts_u4_addr(dst + 32768 + n * 4096, lds_u4_addr(oh_u + slot * 128u));
            } else if (u16 == 8) {
              // live indicator (bf16, K-major [16 x keys], row 0): 1.0 for a visible key, 0 for a padding row
              asm volatile("st.shared.u16 [%0], %1;" ::"r"(sbase + 40960 + (hw + n * 32) * 2), "h"(static_cast<uint16_t>(live ? 0x3f80 : 0)) : "memory");
            }
            if (n + 1 < nkeys) issue(kbase, rc[1], kk);
            blend_store(vv, rc[n], kvb_te + tjo + 256, dst + 16384 + n * 4096, live);
            if (n + 1 < nkeys) issue(vbase, rc[1], vv);
            if (gt == 0) FA_TS(1, ts_i++, 5 + n);                  // key n of the half-warp stored
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&kv_full[stage]);
        if (gt == 0) FA_TS(1, ts_i++, 3);                           // tile delivered
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HMVIT_FA2_MISC_REGS));   // one instruction for the whole warpgroup
    if (warp == Cfg::MMA_WARP + 1) {
      // =========================================== LOADER ===========================================
      // One bulk copy per tile: the tile's (up to) 64 key records, 1 KB, into the record ring; and an L2 prefetch of the
      // query rows two items ahead (256 B of this head group per row = two 128-byte lines).
      const uint32_t rec_u = smem_u32(smem + Cfg::OFF_REC);
      const uint32_t recs_per_item = static_cast<uint32_t>(p.L) * kS;
      uint32_t tcnt = 0;
      int kit = 0;
      for (int it = item0; it < n_items; it += item_step, ++kit) {
        const int it2 = it + 2 * item_step;
        if (it2 < n_items) {
          int a_, g_; item_ag(kit + 2, it2, a_, g_);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int e = lane + 32 * k;                           // (row, half)
            int r, c; group_token(p.kind, g_ / GX, g_ - (g_ / GX) * GX, e >> 1, p.H, p.W, r, c);
            const __nv_bfloat16* rowp = p.q + (static_cast<size_t>(a_) * N + r * p.W + c) * kC + hgc * 128 + (e & 1) * 64;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp) : "memory");
          }
        }
        const int nv = item_nv(kit, it);
        const int ntiles = (nv + kS - 1) >> 6;
        if (ntiles == 0) continue;
        int a, grp; item_ag(kit, it, a, grp);
        const KeyRec* src = fp.rec + (static_cast<size_t>(a) * G + grp) * recs_per_item;
        for (int t = 0; t < ntiles; ++t, ++tcnt) {
          const uint32_t rst = tcnt & (Cfg::REC_STAGES - 1);
          const uint32_t bytes = static_cast<uint32_t>(min(kS, nv - t * kS)) * 16u;
          mbar_wait_sleepy(&rec_empty[rst], ((tcnt / Cfg::REC_STAGES) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(&rec_full[rst], bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(rec_u + rst * 1024), "l"(src + t * kS), "r"(bytes), "r"(smem_u32(&rec_full[rst])) : "memory");
          }
          __syncwarp();
        }
      }
    }
    if (warp == Cfg::MMA_WARP || warp == Cfg::MMA_WARP + 2) {
      // =========================================== MMA ===========================================
      // One MMA warp per head pair (a single warp serving both pairs' Q K^T and P V reacted to each event in 0.5-1 k cycles:
      // the softmax warpgroups waited 2-3 k cycles for an item's final accumulator and for the next item's first S).  Two
      // cursors over the CTA's tile sequence: Q K^T runs at most two tiles ahead of P V (two S buffers).  A ring stage is
      // released by BOTH warps' commits (kv_empty counts 2).
      const int pr = (warp - Cfg::MMA_WARP) >> 1;
      constexpr uint32_t idesc_pv = umma_idesc(1u, 128, 64) | (1u << 16);        // bf16, B (V tile) MN-major
      constexpr uint32_t idesc_l = umma_idesc(1u, 128, 16);                      // bf16, B (live indicator) K-major
      const uint32_t ring_u = smem_u32(sRing), aug_u = smem_u32(smem + Cfg::OFF_AUG) + pr * 16384;
      const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
      const uint32_t tS = tmu + Cfg::TM_S + pr * 128, tQ = tmu + Cfg::TM_Q + pr * 32;
      const uint32_t tD = tmu + Cfg::TM_D + pr * 64, tL = tmu + Cfg::TM_L + pr * 16;
      // S[buf] = Qbd K^T + Bias OneHot^T over the tile's nk key rows
      auto issue_qk = [&](uint32_t stage, uint32_t buf, int nk) {
        if (HMVIT_FA_DBG & 4) return;
        const uint32_t id_bf = umma_idesc(1u, 128, 0) | (static_cast<uint32_t>(nk >> 3) << 17);
        const uint32_t id_h = umma_idesc(0u, 128, 0) | (static_cast<uint32_t>(nk >> 3) << 17);
        const uint32_t sb = ring_u + stage * Cfg::STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ts_bf16(tS + buf * 64, tQ + ks * 8, umma_desc_sw128(sb + pr * 8192 + ks * 32), id_bf, ks != 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss<2>(tS + buf * 64, umma_desc_sw128(aug_u + ks * 32), umma_desc_sw128(sb + 32768 + ks * 32), id_h, 1u);
      };
      auto issue_pv = [&](uint32_t stage, uint32_t buf, bool first, int nk) {     // D (+)= P V, l (+)= P live^T
        if (HMVIT_FA_DBG & 4) return;
        const uint32_t sb = ring_u + stage * Cfg::STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks * 16 < nk)
            umma_ts_bf16(tD, tS + buf * 64 + ks * 8, umma_desc_sw128_mn(sb + 16384 + pr * 8192 + ks * 2048), idesc_pv,
                         (!first || ks != 0) ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          if (ks * 16 < nk)
            umma_ts_bf16(tL, tS + buf * 64 + ks * 8, umma_desc_sw128(sb + 40960 + ks * 32), idesc_l, (!first || ks != 0) ? 1u : 0u);
      };
      auto advance = [&](Fa2Cur& c) -> bool {                      // next tile; true when it starts a new item
        if ((c.t + 1) * kS < c.nv) { ++c.t; return false; }
        c.t = 0; c.it += item_step; ++c.kit;
        seek(c.kit, c.it, c.nv);
        return true;
      };
      Fa2Cur cq, cp;
      uint32_t gq = 0, gp = 0, oq = 0, op = 0;                     // tile counters, item ordinals
      cq.kit = 0; cq.it = item0; cq.t = 0; cq.nv = 0;
      seek(cq.kit, cq.it, cq.nv);
      cp = cq;
      int ts_i = 0; (void)ts_i;
      while (cp.it < n_items) {
        // ---- Q K^T of the next tile: the tile is in the ring, its S buffer is free (two tiles ahead at most), and for an
        // item's first tile the pair's Q is in TMEM ----
        const bool can_qk = cq.it < n_items && gq < gp + HMVIT_FA2_AHEAD, can_pv = gp < gq;
        if (can_qk) {
          const uint32_t stage = gq % Cfg::STAGES, ph = (gq / Cfg::STAGES) & 1u;
          bool ok;
          if (!can_pv) {                                           // nothing else to do: sleep on the barriers
            mbar_wait_sleepy(&kv_full[stage], ph);
            if (cq.t == 0) mbar_wait_sleepy(&q_full[pr], oq & 1u);
            ok = true;
          } else {
            ok = mbar_test_wait(&kv_full[stage], ph) && (cq.t != 0 || mbar_test_wait(&q_full[pr], oq & 1u));
            ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;    // (lane 0's view, so that the whole warp takes the same path)
          }
          if (ok) {
            tc_fence_after();
            const int nk = (min(kS, cq.nv - cq.t * kS) + 15) & ~15;
            if (elect_one()) {
              issue_qk(stage, gq & 1u, nk);
              umma_commit(&s_full[pr * 2 + (gq & 1u)]);
            }
            __syncwarp();
            if (advance(cq)) ++oq;
            ++gq;
          }
        }
        // ---- P V of the oldest tile with a computed S: P is there; for an item's first tile the previous item's accumulator
        // has been read ----
        if (gp < gq) {
          const uint32_t buf = gp & 1u;
          const bool first = cp.t == 0, last = (cp.t + 1) * kS >= cp.nv;
          const bool must = !(cq.it < n_items && gq < gp + HMVIT_FA2_AHEAD);    // no Q K^T possible: sleep on the barriers
          bool ok;
          if (must) {
            mbar_wait_sleepy(&p_full[pr * 2 + buf], (gp >> 1) & 1u);
            if (first) mbar_wait_sleepy(&d_free[pr], (op & 1u) ^ 1u);
            ok = true;
          } else {
            ok = mbar_test_wait(&p_full[pr * 2 + buf], (gp >> 1) & 1u) && (!first || mbar_test_wait(&d_free[pr], (op & 1u) ^ 1u));
            ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
          }
          if (ok) {
            tc_fence_after();
            if (lane == 0) FA_TS(2 + pr, ts_i++, 4);               // P there
            const uint32_t stage = gp % Cfg::STAGES;
            const int nk = (min(kS, cp.nv - cp.t * kS) + 15) & ~15;
            if (elect_one()) {
              issue_pv(stage, buf, first, nk);
              umma_commit(&pv_done[pr]);
              if (last) umma_commit(&d_full[pr]);
              umma_commit(&kv_empty[stage]);                        // (one of the two arrivals that release the stage)
            }
            __syncwarp();
            if (advance(cp)) ++op;
            ++gp;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::MMA_WARP) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TM_COLS>(tm);
  }
}

}  // namespace hmvit
