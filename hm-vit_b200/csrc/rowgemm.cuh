// Typed row-GEMM for the HM-ViT fusion path (sm_100a: TMA-fed tcgen05.mma, accumulators in TMEM).
//
//   out[rows of one agent, Ntot] = epilogue( prologue(A rows)[128 x 256] * W_type^T [256 x Ntot] )
//
// One CTA = one 128-token tile of one agent slot (b, l).  The agent's modality type (camera/LiDAR)
// selects the weight matrix (two TMA tensor maps), LayerNorm affine and bias: this is the "grouped"
// part -- tiles of both types run in the same launch.
//   * A (128 x 256, bf16 or tf32) is produced ONCE by the prologue (typed LayerNorm / cast / row copy)
//     straight into the UMMA canonical K-major SWIZZLE_128B layout in shared memory and stays
//     resident for every N-chunk;
//   * W streams through a 16 KB-stage TMA ring (one stage = 128 output channels x 128 B of K);
//   * each 128-column N-chunk accumulates in one of two TMEM buffers, so the epilogue of chunk c
//     overlaps the MMAs of chunk c+1.
// Warp roles: warps 0-3 prologue + epilogue (thread == token row == TMEM lane), warp 4 TMA producer
// (+ TMEM allocator), warp 5 MMA issuer.
//
// Replaces, per call site, the reference's per-(b, agent) nn.Linear launches:
//   QKV  -- HeteroAttention.to_qkv            hetero_fusion.py:111-140 (+ HeteroLayerNorm base_transformer.py:171-177)
//   OUT  -- HeteroAttention.to_out + residual hetero_fusion.py:142-152, 399
//   FFN  -- HeteroPreNormResidual(HeteroFeedForward) base_transformer.py:129-136,180-192
//   HEAD -- HeteroFusion.mlp_head             bevformer_point_pillar_hetero.py:47-48
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace hmvit {

enum : int { PRO_CM_LN = 0, PRO_CM_CAST = 1, PRO_ROWS_BF16 = 2 };
enum : int { EPI_ROWS_BF16 = 0, EPI_CM_RESID = 1, EPI_CM_GELU = 2, EPI_CM_STORE = 3 };

struct RowGemmParams {
  int B, L, N;                 // scenes, agent slots per scene, tokens per agent (H*W)
  int n_chunks;                // Ntot / 128
  const int* mode;             // [B*L] 0 = camera, 1 = lidar
  const int* record_len;       // [B]
  int tile_ego_only;           // 1: only slot 0 of every scene is processed
  int qkv_select;              // 1: chunks are {Q, K|te=0, K|te=1, V|te=0, V|te=1} x 2, enabled per scene
  int qkv_ego_only;            // with qkv_select: Q only for slot 0, K/V only for te = type(slot 0)
  // A operand
  const float* a_cm;           // PRO_CM_*: [B*L][256][N] fp32 (channel-major, i.e. the NCHW module layout)
  const __nv_bfloat16* a_rows; // PRO_ROWS_BF16: [B*L*N][256]
  const float* ln_gamma;       // [2][256]
  const float* ln_beta;        // [2][256]
  float ln_eps;
  const float2* ln_stats;      // PRO_CM_LN, optional: [B*L][N] (mean, rstd) of every row of A (skips the statistics pass)
  // epilogue
  const float* bias;           // [2][Ntot]
  __nv_bfloat16* out_rows;     // EPI_ROWS_BF16: [n_chunks/2][B*L*N][256]
  float* out_cm;               // EPI_CM_*: [B*out_L][256][N]
  const float* resid_cm;       // EPI_CM_RESID: [B*L][256][N]
  int out_L;                   // slots per scene of out_cm (L, or 1 for the head output)
};

template <int ES>
struct RowGemmCfg {
  static constexpr int BM = 128, BN = 128, K = 256;
  static constexpr int CHUNK_BYTES = BM * 128;             // 16 KB: 128 rows x 128 B of K
  static constexpr int K_PER_CHUNK = 128 / ES;             // 64 bf16 / 32 tf32
  static constexpr int NCHA = K * ES / 128;                // A chunks: 4 (bf16) / 8 (tf32)
  static constexpr int NS = (ES == 2) ? 2 : 4;             // B stages (bf16: 2 so that two CTAs fit one SM)
  static constexpr int A_BYTES = NCHA * CHUNK_BYTES;
  static constexpr int B_BYTES = NS * CHUNK_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = A_BYTES + B_BYTES + BAR_BYTES + 1024;   // + alignment slack
  static constexpr int THREADS = 192;
  static constexpr uint32_t TMEM_COLS = 256;               // 2 accumulator buffers x 128 columns
};

template <int ES, int PRO, int EPI>
__global__ void __launch_bounds__(192, (ES == 2) ? 2 : 1)
rowgemm_kernel(const __grid_constant__ CUtensorMap tmap0, const __grid_constant__ CUtensorMap tmap1,
               const RowGemmParams p) {
  using Cfg = RowGemmCfg<ES>;
  const int a = blockIdx.y;
  const int b = a / p.L, l = a - b * p.L;
  const int nrec = min(p.record_len[b], p.L);
  if (l >= nrec || (p.tile_ego_only && l != 0)) return;
  const int type = p.mode[a] != 0 ? 1 : 0;
  const int tok0 = blockIdx.x * Cfg::BM;

  // which N-chunks this tile computes
  uint32_t chunk_mask = (p.n_chunks >= 32) ? 0xffffffffu : ((1u << p.n_chunks) - 1u);
  if (p.qkv_select) {
    uint32_t te_mask = 0;
    if (p.qkv_ego_only) te_mask = 1u << (p.mode[b * p.L] != 0 ? 1 : 0);
    else for (int j = 0; j < nrec; ++j) te_mask |= 1u << (p.mode[b * p.L + j] != 0 ? 1 : 0);
    uint32_t m = 0;
    if (!p.qkv_ego_only || l == 0) m |= 0x3u;                 // Q
    if (te_mask & 1u) m |= (0x3u << 2) | (0x3u << 6);         // K|te=0, V|te=0
    if (te_mask & 2u) m |= (0x3u << 4) | (0x3u << 8);         // K|te=1, V|te=1
    chunk_mask &= m;
  }

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::A_BYTES + Cfg::B_BYTES);
  uint64_t* b_full = bars;                 // [NS]
  uint64_t* b_empty = bars + Cfg::NS;      // [NS]
  uint64_t* acc_full = bars + 2 * Cfg::NS; // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint64_t* a_full = acc_empty + 2;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const CUtensorMap* tmap = type ? &tmap1 : &tmap0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 128); }
    mbar_init(a_full, 128);
    fence_mbar_init();
  }
  if (warp == 4) {
    if (lane == 0) tma_prefetch_desc(tmap);
    tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ================= prologue: fill the A tile =================
    const int row = threadIdx.x;                       // 0..127
    if constexpr (PRO == PRO_ROWS_BF16) {
      static_assert(ES == 2 || PRO != PRO_ROWS_BF16, "row copy prologue is bf16 only");
      // warp-per-row, lane == 16-byte unit of the 512-byte row
      const uint4* src = reinterpret_cast<const uint4*>(p.a_rows) + (static_cast<size_t>(a) * p.N + tok0) * 32;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int r = warp * 32 + rr;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (tok0 + r < p.N) v = __ldg(src + static_cast<size_t>(r) * 32 + lane);
        *reinterpret_cast<uint4*>(sA + (lane >> 3) * Cfg::CHUNK_BYTES + sw128_offset(r, lane & 7)) = v;
      }
    } else {
      const int tok = tok0 + row;
      const bool valid = tok < p.N;
      const float* src = p.a_cm + static_cast<size_t>(a) * kC * p.N + (valid ? tok : 0);
      {
        // L2 prefetch of the tile's 256 channel rows (512 B = four 128-byte lines each): the channel batches below are
        // dependent rounds of 32 loads per thread, eight exposed HBM round trips per tile without it (the fp32 variants run
        // one CTA per SM, so nothing else hides them)
        const float* tb = p.a_cm + static_cast<size_t>(a) * kC * p.N + tok0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = i * 128 + row, ch = e >> 2, seg = (e & 3) * 32;
          if (tok0 + seg < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(tb + static_cast<size_t>(ch) * p.N + seg));
        }
      }
      float mean = 0.f, rstd = 1.f;
      if (PRO == PRO_CM_LN && p.ln_stats != nullptr) {
        if (valid) { const float2 st = __ldg(p.ln_stats + static_cast<size_t>(a) * p.N + tok); mean = st.x; rstd = st.y; }
      } else if constexpr (PRO == PRO_CM_LN) {
        // shifted single pass statistics (shift = first channel) -- biased variance like nn.LayerNorm
        const float s0 = valid ? __ldg(src) : 0.f;
        float sum = 0.f, sq = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < kC; c0 += 32) {
          float xv[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) xv[e] = __ldg(src + (c0 + e) * p.N);
#pragma unroll
          for (int e = 0; e < 32; ++e) { const float d = xv[e] - s0; sum += d; sq += d * d; }
        }
        const float md = sum * (1.0f / kC);
        mean = s0 + md;
        const float var = fmaxf(sq * (1.0f / kC) - md * md, 0.f);
        rstd = rsqrtf(var + p.ln_eps);
      }
      const bool affine = p.ln_gamma != nullptr;
      const float* gam = affine ? p.ln_gamma + type * kC : nullptr;
      const float* bet = affine ? p.ln_beta + type * kC : nullptr;
      constexpr int EPU = 16 / ES;                      // elements per 16-byte unit
      constexpr int UPB = 32 / EPU;                     // units per batch of 32 channels
#pragma unroll 1
      for (int c0 = 0; c0 < kC; c0 += 32) {
        float xv[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) xv[e] = __ldg(src + (c0 + e) * p.N);
        if constexpr (PRO == PRO_CM_LN) {
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float z = (xv[e] - mean) * rstd;
            xv[e] = affine ? z * __ldg(gam + c0 + e) + __ldg(bet + c0 + e) : z;
          }
        }
        if (!valid) {
#pragma unroll
          for (int e = 0; e < 32; ++e) xv[e] = 0.f;
        }
#pragma unroll
        for (int uu = 0; uu < UPB; ++uu) {
          const float* v = xv + uu * EPU;
          const int u = c0 / EPU + uu;
          uint4 pk;
          if constexpr (ES == 2) {
            pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
            pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
          } else {
            pk.x = __float_as_uint(tf32_rn(v[0])); pk.y = __float_as_uint(tf32_rn(v[1]));
            pk.z = __float_as_uint(tf32_rn(v[2])); pk.w = __float_as_uint(tf32_rn(v[3]));
          }
          *reinterpret_cast<uint4*>(sA + (u >> 3) * Cfg::CHUNK_BYTES + sw128_offset(row, u & 7)) = pk;
        }
      }
    }
    fence_proxy_async_smem();
    mbar_arrive(a_full);

    // ================= epilogue =================
    const int tok = tok0 + row;
    const bool valid = tok < p.N;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    const int Ntot = p.n_chunks * Cfg::BN;
    int ci = 0;
    for (int c = 0; c < p.n_chunks; ++c) {
      if (!((chunk_mask >> c) & 1u)) continue;
      const int buf = ci & 1;
      mbar_wait(&acc_full[buf], (ci >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tmem_ld32(tmem_base + lane_base + buf * Cfg::BN + q * 32, r);
        tmem_ld_wait();
        const int n0 = c * Cfg::BN + q * 32;
        const float* bias = p.bias + type * Ntot + n0;
        if constexpr (EPI == EPI_ROWS_BF16) {
          if (valid) {
            __nv_bfloat16* dst = p.out_rows + (static_cast<size_t>(c >> 1) * p.B * p.L * p.N + static_cast<size_t>(a) * p.N + tok) * kC +
                                 (c & 1) * Cfg::BN + q * 32;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + k)), b1 = __ldg(reinterpret_cast<const float4*>(bias + k + 4));
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(r[k + 0]) + b0.x, __uint_as_float(r[k + 1]) + b0.y);
              pk.y = pack_bf16x2(__uint_as_float(r[k + 2]) + b0.z, __uint_as_float(r[k + 3]) + b0.w);
              pk.z = pack_bf16x2(__uint_as_float(r[k + 4]) + b1.x, __uint_as_float(r[k + 5]) + b1.y);
              pk.w = pack_bf16x2(__uint_as_float(r[k + 6]) + b1.z, __uint_as_float(r[k + 7]) + b1.w);
              *reinterpret_cast<uint4*>(dst + k) = pk;
            }
          }
        } else {
          if (valid) {
            const int lo = (p.out_L == p.L) ? l : 0;
            float* dst = p.out_cm + (static_cast<size_t>(b) * p.out_L + lo) * kC * p.N + static_cast<size_t>(n0) * p.N + tok;
            const float* res = nullptr;
            if constexpr (EPI == EPI_CM_RESID) res = p.resid_cm + static_cast<size_t>(a) * kC * p.N + static_cast<size_t>(n0) * p.N + tok;
            float rv[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) rv[k] = (EPI == EPI_CM_RESID) ? res[k * p.N] : 0.f;     // 32-bit offsets: 256 N < 2^31
            float bq[32];
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + k));
              bq[k] = b4.x; bq[k + 1] = b4.y; bq[k + 2] = b4.z; bq[k + 3] = b4.w;
            }
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              float v = __uint_as_float(r[k]) + bq[k] + rv[k];
              if constexpr (EPI == EPI_CM_GELU) v = tf32_rn(gelu_erf(v));
              dst[k * p.N] = v;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);
      ++ci;
    }
  } else if (warp == 4) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        if (!((chunk_mask >> c) & 1u)) continue;
        for (int kc = 0; kc < Cfg::NCHA; ++kc, ++it) {
          const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1u;
          mbar_wait(&b_empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&b_full[s], Cfg::CHUNK_BYTES);
          tma_load_2d(sB + s * Cfg::CHUNK_BYTES, tmap, &b_full[s], kc * Cfg::K_PER_CHUNK, c * Cfg::BN);
        }
      }
    }
  } else {
    // ================= MMA issuer =================
    // whole warp converged, one elected lane issues (operands stay in uniform registers)
    {
      constexpr uint32_t idesc = umma_idesc(ES == 2 ? 1u : 2u, Cfg::BM, Cfg::BN);
      mbar_wait(a_full, 0);
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      const uint32_t tmu = __shfl_sync(0xffffffffu, tmem_base, 0);
      uint32_t it = 0;
      int ci = 0;
      for (int c = 0; c < p.n_chunks; ++c) {
        if (!((chunk_mask >> c) & 1u)) continue;
        const int buf = ci & 1;
        mbar_wait(&acc_empty[buf], ((ci >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmu + buf * Cfg::BN;
        for (int kc = 0; kc < Cfg::NCHA; ++kc, ++it) {
          const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1u;
          mbar_wait(&b_full[s], ph);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t ad = umma_desc_sw128(a_base + kc * Cfg::CHUNK_BYTES + ks * 32);
              const uint64_t bd = umma_desc_sw128(b_base + s * Cfg::CHUNK_BYTES + ks * 32);
              umma_ss<ES>(d_tmem, ad, bd, idesc, (kc | ks) != 0 ? 1u : 0u);
            }
            umma_commit(&b_empty[s]);
            if (kc == Cfg::NCHA - 1) umma_commit(&acc_full[buf]);
          }
          __syncwarp();
        }
        ++ci;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace hmvit
