// Input gradient of the fused typed Q | K' | V' projection (the adjoint of qkv_kernel's GEMM):
//
//   out[a][c][tok] = sum over the 5 planes p (Q, K'|te=0, K'|te=1, V'|te=0, V'|te=1) and k of
//                    dcat[p][a*N + tok][k] * W_cat[type(a)][p*256 + k][c]
//
// i.e. what autograd computes for the input of HeteroAttention.to_qkv (hetero_fusion.py:111-132) with the folded
// weights of DESIGN.md section 2.  Until this kernel the backward ran it as five K = 256 row-GEMMs accumulating through
// the fp32 output (5 launches, 4 extra read-modify-write passes over a 346 MB tensor per stage); this is ONE K = 1280
// GEMM whose accumulator stays in tensor memory:
//   * persistent CTAs over the (active agent, 128-token tile) list, two 256-column accumulators in tensor memory so the
//     epilogue of a tile overlaps the MMAs of the next;
//   * everything is fed by TMA: dcat rows (bf16 [5*R][256], box 128 tokens x 64 k) and the transposed weights
//     (bf16 [5*256 rows (p, c)][256 k], box 256 x 64 k), both K-major SWIZZLE_128B, 4-stage ring of 48 KB stages;
//   * MMA 128 x 256 x 16 (kind::f16, bf16 operands, fp32 accumulate), 80 per tile;
//   * epilogue: thread = token, channel-major fp32 stores (coalesced across the warp).
// Algorithmic bytes per stage: 5 * 173 MB (dcat) + 346 MB (out) -- HBM-bound.
#pragma once
#include "bwd.cuh"

namespace hmvit {

struct DgradCatParams {
  int L, N, n_agents;
  long long R;                    // rows of one plane = B*L*N
  const int* mode;
  const int* record_len;
  float* out;                     // cm fp32 [B*L][256][N]
};

struct DgradCatCfg {
  static constexpr int BM = 128;                    // tokens per tile
  static constexpr int NS = 4;
  static constexpr int A_BYTES = 128 * 128;         // 128 tokens x 64 k, bf16
  static constexpr int B_BYTES = 256 * 128;         // 256 channels x 64 k, bf16
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NCHUNK = 20;                 // 5 planes x 4 k-chunks
  static constexpr int THREADS = 192;
  static constexpr int MAX_AGENTS = 2048;
  static constexpr int OFF_BARS = NS * STAGE_BYTES;
  static constexpr int OFF_LIST = OFF_BARS + 128;
  static constexpr int SMEM_BYTES = 1024 + OFF_LIST + MAX_AGENTS * 2;
};

__global__ void __launch_bounds__(DgradCatCfg::THREADS, 1)
dgrad_cat_kernel(const __grid_constant__ CUtensorMap a_map, const __grid_constant__ CUtensorMap w_map0,
                 const __grid_constant__ CUtensorMap w_map1, const DgradCatParams p) {
  using Cfg = DgradCatCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* full = bars;                      // [NS]
  uint64_t* empty = bars + Cfg::NS;           // [NS]
  uint64_t* acc_full = bars + 2 * Cfg::NS;    // [2]
  uint64_t* acc_empty = acc_full + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* s_nact = reinterpret_cast<int*>(tmem_slot + 1);
  uint16_t* sList = reinterpret_cast<uint16_t*>(smem + Cfg::OFF_LIST);   // active agents: index | type << 15

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
    fence_mbar_init();
  }
  if (warp == 0) {
    int n = 0;
    for (int a0 = 0; a0 < p.n_agents; a0 += 32) {
      const int a = a0 + lane;
      const bool ok = a < p.n_agents && agent_active(a, p.L, p.record_len, 0);
      const uint32_t bal = __ballot_sync(0xffffffffu, ok);
      if (ok) sList[n + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint16_t>(a | ((p.mode[a] != 0 ? 1 : 0) << 15));
      n += __popc(bal);
    }
    if (lane == 0) *s_nact = n;
  }
  if (warp == 4) {
    if (lane == 0) { tma_prefetch_desc(&a_map); tma_prefetch_desc(&w_map0); tma_prefetch_desc(&w_map1); }
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const int TPA = (p.N + Cfg::BM - 1) / Cfg::BM;
  const int T = *s_nact * TPA;

  if (warp < 4) {
    // ======================= epilogue: tensor memory -> channel-major fp32 =======================
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t j = 0;
    for (int t = blockIdx.x; t < T; t += gridDim.x, ++j) {
      const int ai = t / TPA, a = sList[ai] & 0x7fff;
      const int tok = (t - ai * TPA) * Cfg::BM + threadIdx.x;
      const uint32_t buf = j & 1u;
      mbar_wait(&acc_full[buf], (j >> 1) & 1u);
      tc_fence_after();
      float* dst = p.out + static_cast<size_t>(a) * kC * p.N + tok;
#pragma unroll 1
      for (int q = 0; q < 8; ++q) {
        uint32_t r[32];
        tmem_ld32(tm + lane_base + buf * 256 + q * 32, r);
        tmem_ld_wait();
        if (tok < p.N) {
#pragma unroll
          for (int k = 0; k < 32; ++k) dst[static_cast<size_t>(q * 32 + k) * p.N] = __uint_as_float(r[k]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  } else if (warp == 4) {
    // ======================= TMA producer =======================
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const int ai = t / TPA, e = sList[ai], a = e & 0x7fff;
        const CUtensorMap* wm = (e >> 15) ? &w_map1 : &w_map0;
        const long long row0 = static_cast<long long>(a) * p.N + (t - ai * TPA) * Cfg::BM;
        for (int c = 0; c < Cfg::NCHUNK; ++c, ++it) {
          const int pl = c >> 2, kc = c & 3;
          const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          tma_load_2d(st, &a_map, &full[s], kc * 64, static_cast<int32_t>(pl * p.R + row0));
          tma_load_2d(st + Cfg::A_BYTES, wm, &full[s], kc * 64, pl * 256);
        }
      }
    }
  } else {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc = umma_idesc(1u, 128, 256);
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t sbase = smem_u32(smem);
    uint32_t it = 0, j = 0;
    for (int t = blockIdx.x; t < T; t += gridDim.x, ++j) {
      const uint32_t buf = j & 1u;
      mbar_wait(&acc_empty[buf], ((j >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int c = 0; c < Cfg::NCHUNK; ++c, ++it) {
        const uint32_t s = it % Cfg::NS, ph = (it / Cfg::NS) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = sbase + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss<2>(tmu + buf * 256, umma_desc_sw128(sa + ks * 32), umma_desc_sw128(sb + ks * 32), idesc, (c | ks) != 0 ? 1u : 0u);
          umma_commit(&empty[s]);
          if (c == Cfg::NCHUNK - 1) umma_commit(&acc_full[buf]);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<512>(tm);
  }
}

}  // namespace hmvit
