// Typed weight gradient on the 5th-generation tensor cores (tcgen05, accumulators in tensor memory).
//
//   dW[type][row0 + m][n] += sum over the tokens of every active agent of that type of  A(tok, m) * B(tok, n)
//
// i.e. the weight gradients autograd produces for the typed Linears of the fusion block
//   HeteroAttention.to_qkv / to_out     hetero_fusion.py:111-152
//   HeteroFeedForward                   base_transformer.py:180-192
// (gradients of the FOLDED weights).  Same contract as wgrad_kernel (bwd.cuh, warp-level
// wmma tiles, 145 TFLOP/s) which stays for shapes this kernel does not take (both operands bf16 rows, N % 64 != 0,
// more than 2048 agent slots) and as the independently written cross-check.  A bf16-rows B operand with a cm A operand
// runs with the operands exchanged (`swapped`: the accumulator then holds dW itself and the flush is strided).
//
// The contraction runs over TOKENS (K = all tokens of a type, ~170 k at the bench shape) and the output is one
// 256 x 256 matrix per type, so the kernel is bound by reading its operands once: 346 MB (fp32 cm) + 173 MB
// (bf16 rows) per call.  Design:
//   * persistent: gridDim.x CTAs split the flat list of (active agent, 64-token tile) evenly; a CTA accumulates
//     its whole range in tensor memory and flushes with fp32 reductions once (and once more when the agent type
//     changes inside its range);
//   * D is kept TRANSPOSED in tensor memory, D'[n][m] (lane = n, column = m): the n-side operand (math B, always
//     cm fp32, optionally LayerNorm-normalised with per-token statistics) is the UMMA A operand, K-major -- the
//     cm layout has the tokens contiguous, which IS K-major; stager warps convert it to bf16 and write the UMMA
//     canonical SWIZZLE_128B layout.  With lane = n a warp's reduction for one m covers 32 consecutive floats of
//     dW[m][:] (coalesced 128-byte reductions);
//   * the m-side operand (math A) is the UMMA B operand: bf16 rows [tok][256] are already MN-major, so TMA drops
//     [64 tok][64 ch] boxes (SWIZZLE_128B) straight into shared memory, no register staging, and the MMAs take N = 64
//     column blocks (the layout the fused attention's V tile uses); a cm m-side operand is staged like the n-side
//     one (K-major, N = 256);
//   * 512 tensor-memory columns: [n half h][256 m] at column h * 256.
#pragma once
#include "bwd.cuh"

namespace hmvit {

struct WgradTcParams {
  int L, N, n_agents;
  const int* mode;
  const int* record_len;
  int ego_only;
  const float* m_cm;              // m-side operand, cm (M_ROWS == false); rows come through the tensor map
  const float* n_cm;              // n-side operand, cm
  const float2* n_stats;          // optional: normalise the n-side operand with per-token (mean, rstd)
  float* dw;                      // [2][dw_rows][256] fp32, accumulated
  int dw_rows, dw_row0;
  int swapped;                    // the caller exchanged the operands (math B is bf16 rows): D' holds dW itself, not its
                                  // transpose -- lane = dW row, column = dW column (strided reductions in the flush)
};

template <bool M_ROWS>
struct WgTc {
  static constexpr int KT = 64;                     // tokens per tile
  static constexpr int NS = 3;                      // ring stages
  static constexpr int OP_BYTES = 256 * 128;        // one operand tile: 256 channels x 64 tokens, bf16
  static constexpr int STAGE_BYTES = 2 * OP_BYTES;
  static constexpr int NW_STAGE = 16;               // stager warps: 8 per register-staged operand, or (M_ROWS) two groups of 8
                                                    // that take alternate tiles (twice the loads in flight)
  static constexpr int NW_TILE = M_ROWS ? 8 : 16;   // stager warps that arrive on a stage's full barrier
  static constexpr int THREADS = (NW_STAGE + 2) * 32;
  static constexpr int MAX_AGENTS = 2048;
  static constexpr int OFF_BARS = NS * STAGE_BYTES;
  static constexpr int OFF_LIST = OFF_BARS + 128;
  static constexpr int SMEM_BYTES = 1024 + OFF_LIST + MAX_AGENTS * 2;
};

template <bool M_ROWS>
__global__ void __launch_bounds__(WgTc<M_ROWS>::THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap m_map, const WgradTcParams p) {
  using Cfg = WgTc<M_ROWS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* full = bars;                    // [NS] operands of the stage are in shared memory
  uint64_t* empty = bars + Cfg::NS;         // [NS] the stage's MMAs have completed
  uint64_t* acc_full = bars + 2 * Cfg::NS;  // accumulator of a segment complete
  uint64_t* acc_empty = acc_full + 1;       // accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
  int* s_nact = reinterpret_cast<int*>(tmem_slot + 1);
  uint16_t* sList = reinterpret_cast<uint16_t*>(smem + Cfg::OFF_LIST);   // active agents: index | type << 15

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::NS; ++s) { mbar_init(&full[s], Cfg::NW_TILE + (M_ROWS ? 1 : 0)); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, Cfg::NW_STAGE);
    fence_mbar_init();
  }
  if (warp == 0) {
    int n = 0;
    for (int a0 = 0; a0 < p.n_agents; a0 += 32) {
      const int a = a0 + lane;
      const bool ok = a < p.n_agents && agent_active(a, p.L, p.record_len, p.ego_only);
      const uint32_t bal = __ballot_sync(0xffffffffu, ok);
      if (ok) sList[n + __popc(bal & ((1u << lane) - 1u))] = static_cast<uint16_t>(a | ((p.mode[a] != 0 ? 1 : 0) << 15));
      n += __popc(bal);
    }
    if (lane == 0) *s_nact = n;
  }
  if (warp == Cfg::NW_STAGE) {
    if (M_ROWS && lane == 0) tma_prefetch_desc(&m_map);
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;
  const int TPA = p.N / Cfg::KT;                                   // tiles per agent
  const long long T = static_cast<long long>(*s_nact) * TPA;
  const int t0 = static_cast<int>(T * blockIdx.x / gridDim.x), t1 = static_cast<int>(T * (blockIdx.x + 1) / gridDim.x);

  if (warp < Cfg::NW_STAGE) {
    // ======================= stagers (cm fp32 -> bf16, K-major SWIZZLE_128B) + accumulator flush =======================
    const bool n_side = M_ROWS || warp < 8;
    const int st = threadIdx.x & 255, u = st & 7, r0 = st >> 3;    // 16-byte unit (8 tokens), first of the thread's 8 rows
    const float* src_base = n_side ? p.n_cm : p.m_cm;
    const float2* stats = n_side ? p.n_stats : nullptr;
    const uint32_t op_off = n_side ? 0u : Cfg::OP_BYTES;

    auto flush = [&](int type, uint32_t seg) {
      mbar_wait(acc_full, seg & 1u);
      tc_fence_after();
      constexpr int RANGE = 512 / (Cfg::NW_STAGE / 4);             // columns per warp
      const int q = warp & 3, col0 = (warp >> 2) * RANGE;
      const int h = col0 >> 8, m0 = col0 & 255;
      const int nl = h * 128 + q * 32 + lane;                      // this thread's n-side channel
      float* dst = p.dw + (static_cast<size_t>(type) * p.dw_rows + p.dw_row0) * kC +
                   (p.swapped ? static_cast<size_t>(nl) * kC + m0 : static_cast<size_t>(m0) * kC + nl);
#pragma unroll 1
      for (int c = 0; c < RANGE; c += 32) {
        uint32_t r[32];
        tmem_ld32(tm + (static_cast<uint32_t>(q * 32) << 16) + col0 + c, r);
        tmem_ld_wait();
        if (p.swapped) {                  // 32 consecutive floats of one dW row per thread: 128-bit vector reductions
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c + k), "f"(__uint_as_float(r[k])),
                         "f"(__uint_as_float(r[k + 1])), "f"(__uint_as_float(r[k + 2])), "f"(__uint_as_float(r[k + 3])) : "memory");
        } else {                          // one dW row per register, the warp's lanes cover 32 consecutive columns
#pragma unroll
          for (int k = 0; k < 32; ++k) atomicAdd(dst + static_cast<size_t>(c + k) * kC, __uint_as_float(r[k]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    };

    int cur_type = -1;
    uint32_t seg = 0, i = 0;
    for (int t = t0; t < t1; ++t, ++i) {
      const int ai = t / TPA, kt = t - ai * TPA;
      const int e = sList[ai], a = e & 0x7fff, type = e >> 15;
      if (cur_type >= 0 && type != cur_type) { flush(cur_type, seg); ++seg; }
      cur_type = type;
      if (M_ROWS && (i & 1u) != static_cast<uint32_t>(warp >> 3)) continue;      // the other group's tile
      const int tok0 = kt * Cfg::KT;
      const float* src = src_base + (static_cast<size_t>(a) * kC + r0) * p.N + tok0 + u * 8;
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* s4 = reinterpret_cast<const float4*>(src + static_cast<size_t>(j) * 32 * p.N);
        v[2 * j] = __ldg(s4);
        v[2 * j + 1] = __ldg(s4 + 1);
      }
      {   // L2 prefetch of this thread's pieces of the next tile it stages (its loads then see L2, not HBM, latency)
        const int tn = t + (M_ROWS ? 2 : 1);
        if (tn < t1) {
          const int ain = tn / TPA, an = sList[ain] & 0x7fff;
          const float* nsrc = src_base + (static_cast<size_t>(an) * kC + r0) * p.N + (tn - ain * TPA) * Cfg::KT + u * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + static_cast<size_t>(j) * 32 * p.N));
        }
      }
      float4 sv[4];
      if (stats != nullptr) {
        const float4* sp = reinterpret_cast<const float4*>(stats + static_cast<size_t>(a) * p.N + tok0 + u * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j) sv[j] = __ldg(sp + j);
      }
      const uint32_t s = i % Cfg::NS, ph = (i / Cfg::NS) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* dst = smem + s * Cfg::STAGE_BYTES + op_off;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 x0 = v[2 * j], x1 = v[2 * j + 1];
        if (stats != nullptr) {
          x0.x = (x0.x - sv[0].x) * sv[0].y; x0.y = (x0.y - sv[0].z) * sv[0].w;
          x0.z = (x0.z - sv[1].x) * sv[1].y; x0.w = (x0.w - sv[1].z) * sv[1].w;
          x1.x = (x1.x - sv[2].x) * sv[2].y; x1.y = (x1.y - sv[2].z) * sv[2].w;
          x1.z = (x1.z - sv[3].x) * sv[3].y; x1.w = (x1.w - sv[3].z) * sv[3].w;
        }
        const uint4 pk = make_uint4(pack_bf16x2(x0.x, x0.y), pack_bf16x2(x0.z, x0.w), pack_bf16x2(x1.x, x1.y), pack_bf16x2(x1.z, x1.w));
        sts_u4(dst + sw128_offset(static_cast<uint32_t>(r0 + 32 * j), static_cast<uint32_t>(u)), pk);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    if (cur_type >= 0) flush(cur_type, seg);
  } else if (warp == Cfg::NW_STAGE) {
    // ======================= TMA producer: bf16 rows of the m-side operand =======================
    if (M_ROWS && lane == 0) {
      uint32_t i = 0;
      for (int t = t0; t < t1; ++t, ++i) {
        const int ai = t / TPA, kt = t - ai * TPA;
        const int a = sList[ai] & 0x7fff;
        const uint32_t s = i % Cfg::NS, ph = (i / Cfg::NS) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], Cfg::OP_BYTES);
        uint8_t* dst = smem + s * Cfg::STAGE_BYTES + Cfg::OP_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) tma_load_2d(dst + j * 8192, &m_map, &full[s], j * 64, a * p.N + kt * Cfg::KT);
      }
    }
  } else {
    // ======================= MMA issuer =======================
    constexpr uint32_t idesc_mn = umma_idesc(1u, 128, 64) | (1u << 16);   // bf16, B MN-major (TMA'd rows)
    constexpr uint32_t idesc_k = umma_idesc(1u, 128, 256);                // bf16, both K-major
    const uint32_t tmu = __shfl_sync(0xffffffffu, tm, 0);
    const uint32_t sbase = smem_u32(smem);
    int cur_type = -1;
    uint32_t seg = 0, i = 0;
    bool first = true;
    for (int t = t0; t < t1; ++t, ++i) {
      const int ai = t / TPA;
      const int type = sList[ai] >> 15;
      if (cur_type >= 0 && type != cur_type) {
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
        mbar_wait(acc_empty, seg & 1u);
        tc_fence_after();
        ++seg;
        first = true;
      }
      cur_type = type;
      const uint32_t s = i % Cfg::NS, ph = (i / Cfg::NS) & 1u;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = sbase + s * Cfg::STAGE_BYTES, sb = sa + Cfg::OP_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t acc = (first && ks == 0) ? 0u : 1u;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t ad = umma_desc_sw128(sa + h * 16384 + ks * 32);
            if constexpr (M_ROWS) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_ss<2>(tmu + h * 256 + j * 64, ad, umma_desc_sw128(sb + j * 8192 + ks * 2048), idesc_mn, acc);
            } else {
              umma_ss<2>(tmu + h * 256, ad, umma_desc_sw128(sb + ks * 32), idesc_k, acc);
            }
          }
        }
        umma_commit(&empty[s]);
      }
      __syncwarp();
      first = false;
    }
    if (cur_type >= 0) {
      if (elect_one()) umma_commit(acc_full);
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == Cfg::NW_STAGE) {
    tc_fence_after();
    tmem_dealloc<512>(tm);
  }
}

}  // namespace hmvit
