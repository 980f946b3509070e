// Typed LayerNorm + Q / K' / V' projection (sm_100a, persistent, warp-specialised tcgen05 GEMM).
//
//   [Q | K'(te=0) | K'(te=1) | V'(te=0) | V'(te=1)] (bf16 rows) = LN_type(x)[128 x 256] * W_type^T [256 x 1280] + b
//
// Replaces HeteroLayerNorm (base_transformer.py:171-177) + HeteroAttention.to_qkv
// (hetero_fusion.py:111-140) + get_hetero_edge_weights (:154-185; relation_att / relation_msg are
// folded into W_k / W_v on the host) for every valid agent of every scene in one launch.
//
// One CTA per SM loops over 128-token tiles.  Warp roles (18 warps; numbering in QkvCfg):
//   8 warps    A producers: typed LayerNorm of the NEXT tile (channel-major fp32 -> packed bf16) written with
//              tcgen05.st into one of two A buffers IN TENSOR MEMORY (thread == token row == TMEM lane, two
//              warps per lane quarter, 128 channels each), overlapped with the current tile's MMAs.  The A
//              operand never touches shared memory: with A in smem every M128 N128 K16 MMA read 8 KB per
//              64 cycles = the whole 128 B/clk shared-memory port, which starved the epilogue's staging
//              (measured: 1.4-2.0 k cycles for 130 instructions; tools/qkv_timeline.py)
//   1 warp     TMA producer: weight stages (128 output channels x 128 B of K), NS-deep ring
//   1 warp     MMA issuer: 16 x tcgen05.mma (A from TMEM, M128 N128 K16) per 128-column chunk, 2 accumulators
//   warps 0-7  epilogue: TMEM -> +bias -> bf16 -> swizzled smem staging -> fully coalesced row stores (a TMA tensor
//              store from the staging tile was measured slower: 4 KB boxes queue behind the weight loads in the
//              SM's TMA unit, the staging buffer stayed busy for ~3 k cycles)
// Only the chunks a scene needs are computed (K'/V' for the ego types present; in the last stage Q
// for slot 0 only).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace hmvit {

#ifdef HMVIT_TS   // timeline instrumentation build: CTA 0 records clock64() per role (tools/qkv_timeline.py)
__device__ unsigned long long g_qkv_ts[3][512];   // [role: 0 epilogue warp 0, 1 MMA issuer, 2 producer warp 0][event]
#define QKV_TS(role, idx) do { if (blockIdx.x == 0 && (idx) < 512) g_qkv_ts[role][idx] = clock64(); } while (0)
#else
#define QKV_TS(role, idx) do { } while (0)
#endif

struct QkvParams {
  int B, L, N;
  const int* mode;             // [B*L]
  const int* record_len;       // [B]
  int ego_only;                // Q only for slot 0, K'/V' only for te = type(slot 0)
  const float* x_cm;           // [B*L][256][N] fp32
  const float* ln_gamma;       // [2][256] or null (affine folded into W / bias on the host)
  const float* ln_beta;        // [2][256] or null
  float ln_eps;
  const float2* stats_in;      // optional [B*L][N] (mean, rstd) per row written by the previous stage, or null
  const float* bias;           // [2][1280]
  __nv_bfloat16* out_rows;     // [5][B*L*N][256]
};

struct QkvCfg {
  static constexpr int BM = 128, BN = 128;
  static constexpr int CHUNK = 16384;
  static constexpr int NCHA = 4;                       // K chunks of 64 channels (one weight stage each)
#ifndef HMVIT_QKV_NS
#define HMVIT_QKV_NS 2
#endif
#ifndef HMVIT_QKV_SUB      // 64-channel K chunks per weight stage: a successful mbarrier wait costs the single MMA-issuing
#define HMVIT_QKV_SUB 4    // thread 200-300 cycles (tools/qkv_timeline.py), so stages are made large and waits few
#endif
#ifndef HMVIT_QKV_ROT
#define HMVIT_QKV_ROT 0
#endif
#ifndef HMVIT_QKV_DBG      // bottleneck-hunting builds only (results are wrong): 1 no stores, 2 no weight TMA, 4 no MMA, 8 no x loads
#define HMVIT_QKV_DBG 0
#endif
  static constexpr int NS = HMVIT_QKV_NS;              // weight ring stages
  static constexpr int SUB = HMVIT_QKV_SUB;            // K chunks per stage
  static constexpr int STAGE = SUB * CHUNK;            // bytes per weight stage
  static_assert(NCHA % SUB == 0, "stage size");
  static constexpr int N_CHUNKS = 10;                  // 1280 / 128
  static constexpr uint32_t A_COLS = kC / 2;           // TMEM columns per A buffer: 128 lanes x 256 bf16, two per column
  static constexpr int EPI_WARPS = 8;                  // 2 per SM sub-partition: (TMEM lane quarter, 64-column half of the chunk)
  static constexpr int STAGE_BYTES = EPI_WARPS * 32 * 128;   // epilogue staging: per warp 32 rows x 128 B (64 columns)
  static constexpr int BIAS_BYTES = 2 * N_CHUNKS * BN * 4; // both types' [1280] biases, read by every epilogue thread
  static constexpr int PART_BYTES = 2 * 128 * 8;       // partial LayerNorm sums exchanged between the two threads of a row
  static constexpr int SMEM_BYTES = NS * STAGE + STAGE_BYTES + BIAS_BYTES + PART_BYTES + 256 + 1024;
#ifndef HMVIT_QKV_PW
#define HMVIT_QKV_PW 8
#endif
  static constexpr int PROD_WARPS = HMVIT_QKV_PW;      // 4: one thread per token row; 8: two threads per row (128 channels each)
  static constexpr int THREADS = (EPI_WARPS + 2 + PROD_WARPS) * 32;   // epilogue | TMA | MMA | A producers
  static constexpr int W_TMA = EPI_WARPS, W_MMA = EPI_WARPS + 1, W_PROD0 = EPI_WARPS + 2;
  static constexpr int NB = 2;                         // TMEM accumulator buffers (128 columns each): MMA runs one chunk ahead
  static constexpr uint32_t TM_ACC = 2 * A_COLS;       // columns: A buffer 0 | A buffer 1 | accumulator 0 | accumulator 1
  static constexpr uint32_t TMEM_COLS = TM_ACC + NB * BN;
  static_assert(TMEM_COLS == 512, "TMEM budget");
};

template <bool kLN>
__global__ void __launch_bounds__(QkvCfg::THREADS, 1)
qkv_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1, const QkvParams p) {
  using Cfg = QkvCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sB = smem;                                   // [NS][CHUNK]
  uint8_t* sStage = sB + Cfg::NS * Cfg::STAGE;          // [8 warps][32 rows][128 B]
  float* sBias = reinterpret_cast<float*>(sStage + Cfg::STAGE_BYTES);   // [2][1280]
  float2* sPart = reinterpret_cast<float2*>(sStage + Cfg::STAGE_BYTES + Cfg::BIAS_BYTES);   // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + Cfg::STAGE_BYTES + Cfg::BIAS_BYTES + Cfg::PART_BYTES);
  uint64_t* b_full = bars;                  // [NS]
  uint64_t* b_empty = b_full + Cfg::NS;     // [NS]
  uint64_t* acc_full = b_empty + Cfg::NS;   // [NB]
  uint64_t* acc_empty = acc_full + Cfg::NB; // [NB]
  uint64_t* a_full = acc_empty + Cfg::NB;   // [2]
  uint64_t* a_empty = a_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < Cfg::NB; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], Cfg::EPI_WARPS * 32); }
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], Cfg::PROD_WARPS * 32); mbar_init(&a_empty[s], 1); }
    fence_mbar_init();
  }
  if (warp == Cfg::W_MMA) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  // biases live in shared memory: a global (L2) load per use stalled the epilogue warps (ncu: long scoreboard)
  for (int e = threadIdx.x; e < 2 * Cfg::N_CHUNKS * Cfg::BN; e += Cfg::THREADS) sBias[e] = __ldg(p.bias + e);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_agent = (p.N + Cfg::BM - 1) / Cfg::BM;
  const int total_tiles = p.B * p.L * tiles_per_agent;

  // tile -> (agent, first token, chunk mask); false for padded agent slots
  auto tile_info = [&](int t, int& a, int& tok0, uint32_t& chunk_mask) -> bool {
    a = t / tiles_per_agent;
    tok0 = (t - a * tiles_per_agent) * Cfg::BM;
    const int b = a / p.L, l = a - b * p.L;
    const int nrec = min(p.record_len[b], p.L);
    if (l >= nrec) return false;
    uint32_t te_mask = 0;
    if (p.ego_only) te_mask = 1u << (p.mode[b * p.L] != 0 ? 1 : 0);
    else for (int j = 0; j < nrec; ++j) te_mask |= 1u << (p.mode[b * p.L + j] != 0 ? 1 : 0);
    uint32_t m = 0;
    if (!p.ego_only || l == 0) m |= 0x3u;                     // Q
    if (te_mask & 1u) m |= (0x3u << 2) | (0x3u << 6);         // K'|te=0, V'|te=0
    if (te_mask & 2u) m |= (0x3u << 4) | (0x3u << 8);         // K'|te=1, V'|te=1
    chunk_mask = m;
    return true;
  };

  // CTAs walk the 10 weight chunks in a rotated order so that the 148 SMs do not all stream the same
  // L2 lines at the same time
  const int rot = HMVIT_QKV_ROT ? static_cast<int>(blockIdx.x % Cfg::N_CHUNKS) : 0;

  if (warp < Cfg::EPI_WARPS) {
    // ============================ epilogue ============================
    // One warp per scheduler could not hide its own ALU / TMEM / shared-memory latencies (the kernel ran at the
    // speed of this instruction stream even with loads, stores and MMAs removed), hence two warps per sub-partition:
    // warp w owns TMEM lanes 32 (w & 3) .. +31 (= tile rows) and the 64-column half (w >> 2) of every chunk.
    const int q4 = warp & 3, chh = warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    uint8_t* stg = sStage + warp * (32 * 128);
    const size_t rows_total = static_cast<size_t>(p.B) * p.L * p.N;
    uint32_t ci = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int a, tok0; uint32_t chunk_mask;
      if (!tile_info(t, a, tok0, chunk_mask)) continue;
      const int type = p.mode[a] != 0 ? 1 : 0;
      for (int cc = 0; cc < Cfg::N_CHUNKS; ++cc) {
        const int c = (cc + rot) % Cfg::N_CHUNKS;
        if (!((chunk_mask >> c) & 1u)) continue;
        const uint32_t buf = ci % Cfg::NB;
        const float* bb = sBias + type * (Cfg::N_CHUNKS * Cfg::BN) + c * Cfg::BN + chh * 64;
        __nv_bfloat16* obase = p.out_rows + (static_cast<size_t>(c >> 1) * rows_total + static_cast<size_t>(a) * p.N + tok0 + q4 * 32) * kC +
                               (c & 1) * Cfg::BN + chh * 64;
        if (threadIdx.x == 0) QKV_TS(0, ci * 5 + 0);
        mbar_wait(&acc_full[buf], (ci / Cfg::NB) & 1u);
        tc_fence_after();
        if (threadIdx.x == 0) QKV_TS(0, ci * 5 + 1);
        uint32_t r0[32], r1[32];
        tmem_ld32(tmem_base + lane_base + Cfg::TM_ACC + buf * Cfg::BN + chh * 64, r0);
        tmem_ld32(tmem_base + lane_base + Cfg::TM_ACC + buf * Cfg::BN + chh * 64 + 32, r1);
        // Shared-memory loads are requested one step ahead of their use (volatile asm keeps the order): under the
        // MMA's operand traffic an LDS takes ~150 cycles, and a load-use pair per step left 8 + 3 of those
        // exposed per chunk (ncu source view: short-scoreboard stalls on the first FADD / STG after each LDS)
        float4 b0 = lds_f4(bb), b1 = lds_f4(bb + 4);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&acc_empty[buf]);               // accumulator columns are in registers: the MMA warp may refill the buffer
        if (threadIdx.x == 0) QKV_TS(0, ci * 5 + 2);
#pragma unroll
        for (int k8 = 0; k8 < 8; ++k8) {
          const uint32_t* r = (k8 < 4) ? (r0 + k8 * 8) : (r1 + (k8 - 4) * 8);
          float4 n0 = b0, n1 = b1;
          if (k8 < 7) { n0 = lds_f4(bb + (k8 + 1) * 8); n1 = lds_f4(bb + (k8 + 1) * 8 + 4); }
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(r[0]) + b0.x, __uint_as_float(r[1]) + b0.y);
          pk.y = pack_bf16x2(__uint_as_float(r[2]) + b0.z, __uint_as_float(r[3]) + b0.w);
          pk.z = pack_bf16x2(__uint_as_float(r[4]) + b1.x, __uint_as_float(r[5]) + b1.y);
          pk.w = pack_bf16x2(__uint_as_float(r[6]) + b1.z, __uint_as_float(r[7]) + b1.w);
          sts_u4(stg + lane * 128 + ((k8 ^ (lane & 7)) << 4), pk);
          b0 = n0; b1 = n1;
        }
        __syncwarp();
        if (threadIdx.x == 0) QKV_TS(0, ci * 5 + 3);
        // coalesced stores: 4 rows x 128 B per instruction; all eight staging reads are in flight before the first store
        {
          uint4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3), u = lane & 7;
            v[it] = lds_u4(stg + rr * 128 + ((u ^ (rr & 7)) << 4));
          }
          if (HMVIT_QKV_DBG & 1) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if (v[it].x == 0x12345678u) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(it * 4 + (lane >> 3)) * kC + (lane & 7) * 8) = v[it];
          } else if (tok0 + Cfg::BM <= p.N) {        // whole tile inside the agent's token range (the common case)
#pragma unroll
            for (int it = 0; it < 8; ++it)
              *reinterpret_cast<uint4*>(obase + static_cast<size_t>(it * 4 + (lane >> 3)) * kC + (lane & 7) * 8) = v[it];
          } else {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + (lane >> 3);
              if (tok0 + q4 * 32 + rr < p.N) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(rr) * kC + (lane & 7) * 8) = v[it];
            }
          }
        }
        __syncwarp();
        if (threadIdx.x == 0) QKV_TS(0, ci * 5 + 4);
        ++ci;
      }
    }
  } else if (warp == Cfg::W_TMA) {
    // ============================ TMA producer (weights) ============================
    if (lane == 0) {
      tma_prefetch_desc(&tmap0); tma_prefetch_desc(&tmap1);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int a, tok0; uint32_t chunk_mask;
        if (!tile_info(t, a, tok0, chunk_mask)) continue;
        const CUtensorMap* tmap = (p.mode[a] != 0) ? &tmap1 : &tmap0;
        for (int cc = 0; cc < Cfg::N_CHUNKS; ++cc) {
          const int c = (cc + rot) % Cfg::N_CHUNKS;
          if (!((chunk_mask >> c) & 1u)) continue;
          for (int kc = 0; kc < Cfg::NCHA; kc += Cfg::SUB, ++it) {
            const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1u;
            mbar_wait(&b_empty[s], ph ^ 1u);
            if ((HMVIT_QKV_DBG & 2) && it >= Cfg::NS) { mbar_arrive(&b_full[s]); continue; }
            mbar_arrive_expect_tx(&b_full[s], Cfg::STAGE);
#pragma unroll
            for (int j = 0; j < Cfg::SUB; ++j)
              tma_load_2d(sB + s * Cfg::STAGE + j * Cfg::CHUNK, tmap, &b_full[s], (kc + j) * 64, c * Cfg::BN);
          }
        }
      }
    }
  } else if (warp == Cfg::W_MMA) {
    // ============================ MMA issuer ============================
    // The whole warp runs the loop (uniform control flow); one elected lane issues the MMAs and commits.
    {
      constexpr uint32_t idesc = umma_idesc(1u, Cfg::BM, Cfg::BN);
      const uint32_t b_base = smem_u32(sB);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t it = 0, ci = 0, ti = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int a, tok0; uint32_t chunk_mask;
        if (!tile_info(t, a, tok0, chunk_mask)) continue;
        const uint32_t ab = ti & 1u;
        mbar_wait(&a_full[ab], (ti >> 1) & 1u);
        tc_fence_after();
        const uint32_t a_tmem = tm + ab * Cfg::A_COLS;
        for (int cc = 0; cc < Cfg::N_CHUNKS; ++cc) {
          const int c = (cc + rot) % Cfg::N_CHUNKS;
          if (!((chunk_mask >> c) & 1u)) continue;
          const uint32_t buf = ci % Cfg::NB;
          if (lane == 0) QKV_TS(1, ci * 3 + 0);
          mbar_wait(&acc_empty[buf], ((ci / Cfg::NB) & 1u) ^ 1u);
          tc_fence_after();
          if (lane == 0) QKV_TS(1, ci * 3 + 1);
          const uint32_t d_tmem = tm + Cfg::TM_ACC + buf * Cfg::BN;
          for (int kc = 0; kc < Cfg::NCHA; kc += Cfg::SUB, ++it) {
            const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1u;
            mbar_wait(&b_full[s], ph);
            tc_fence_after();
            if (elect_one()) {
              if (!(HMVIT_QKV_DBG & 4)) {
#pragma unroll
                for (int j = 0; j < Cfg::SUB; ++j)
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks)
                    umma_ts_bf16(d_tmem, a_tmem + (kc + j) * 32 + ks * 8,
                                 umma_desc_sw128(b_base + s * Cfg::STAGE + j * Cfg::CHUNK + ks * 32), idesc, (kc | j | ks) != 0 ? 1u : 0u);
              }
              umma_commit(&b_empty[s]);
              if (kc + Cfg::SUB >= Cfg::NCHA) umma_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
          if (lane == 0) QKV_TS(1, ci * 3 + 2);
          ++ci;
        }
        if (elect_one()) umma_commit(&a_empty[ab]);               // every MMA that reads this A buffer has retired
        __syncwarp();
        ++ti;
      }
    }
  } else {
    // ============================ A producers: typed LayerNorm ============================
    // PROD_WARPS == 8: two threads per token row, each owns 128 channels -> twice the loads in flight per SM.
    // A warp can only touch the TMEM lane quarter (warp % 4), which fixes the rows it produces.
    const int pidx = threadIdx.x - Cfg::W_PROD0 * 32;
    const int row = (warp & 3) * 32 + lane, half = (warp - Cfg::W_PROD0) >> 2;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    constexpr int CPT = kC / (Cfg::PROD_WARPS / 4);    // channels per thread: 256 or 128
    const int c_lo = half * CPT;
    uint32_t ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int a, tok0; uint32_t chunk_mask;
      if (!tile_info(t, a, tok0, chunk_mask)) continue;
      const int type = p.mode[a] != 0 ? 1 : 0;
      const uint32_t ab = ti & 1u;
      if (pidx == 0) QKV_TS(2, ti * 4 + 0);
      const int tok = tok0 + row;
      const bool valid = tok < p.N;
      const float* src = p.x_cm + static_cast<size_t>(a) * kC * p.N + (valid ? tok : 0);
      float mean = 0.f, rstd = 1.f;
      if constexpr (kLN) {
        if (p.stats_in != nullptr) {
          if (valid) { const float2 st = __ldg(p.stats_in + static_cast<size_t>(a) * p.N + tok); mean = st.x; rstd = st.y; }
        } else {
          const float s0 = valid ? __ldg(src) : 0.f;       // common shift of both halves
          float sum = 0.f, sq = 0.f;
#pragma unroll 1
          for (int c0 = c_lo; c0 < c_lo + CPT; c0 += 64) {
            float xv[64];
#pragma unroll
            for (int e = 0; e < 64; ++e) xv[e] = __ldg(src + (c0 + e) * p.N);       // row clamped to a valid token: always in bounds
#pragma unroll
            for (int e = 0; e < 64; ++e) { const float d = xv[e] - s0; sum += d; sq += d * d; }
          }
          if constexpr (Cfg::PROD_WARPS == 8) {
            sPart[half * 128 + row] = make_float2(sum, sq);
            named_bar_sync(2, Cfg::PROD_WARPS * 32);
            const float2 o = sPart[(half ^ 1) * 128 + row];
            sum += o.x; sq += o.y;
            named_bar_sync(2, Cfg::PROD_WARPS * 32);         // sPart is rewritten by the next tile
          }
          const float md = sum * (1.0f / kC);
          mean = s0 + md;
          rstd = rsqrtf(fmaxf(sq * (1.0f / kC) - md * md, 0.f) + p.ln_eps);
        }
      }
      const bool affine = kLN && p.ln_gamma != nullptr;
      const float* gam = affine ? p.ln_gamma + type * kC : nullptr;
      const float* bet = affine ? p.ln_beta + type * kC : nullptr;
      const float nmr = -mean * rstd;
      const uint32_t dstA = tmem_base + lane_base + ab * Cfg::A_COLS;
      bool waited = false;
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_lo + CPT; c0 += 64) {
        float xv[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) xv[e] = !(HMVIT_QKV_DBG & 8) ? __ldg(src + (c0 + e) * p.N) : 0.f;   // 32-bit offsets: 256 N < 2^31
        if constexpr (kLN) {
          if (affine) {
#pragma unroll
            for (int e = 0; e < 64; ++e) xv[e] = fmaf(xv[e], rstd, nmr) * __ldg(gam + c0 + e) + __ldg(bet + c0 + e);
          } else {
#pragma unroll
            for (int e = 0; e < 64; ++e) xv[e] = fmaf(xv[e], rstd, nmr);
          }
        }
        if (!valid) {
#pragma unroll
          for (int e = 0; e < 64; ++e) xv[e] = 0.f;
        }
        if (!waited) {
          if (pidx == 0) QKV_TS(2, ti * 4 + 1);
          mbar_wait(&a_empty[ab], ((ti >> 1) & 1u) ^ 1u); waited = true;   // loads above overlap the wait
          tc_fence_after();
          if (pidx == 0) QKV_TS(2, ti * 4 + 2);
        }
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) pk[e] = pack_bf16x2(xv[2 * e], xv[2 * e + 1]);
        tmem_st32(dstA + c0 / 2, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&a_full[ab]);
      if (pidx == 0) QKV_TS(2, ti * 4 + 3);
      ++ti;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::W_MMA) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace hmvit
