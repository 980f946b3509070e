// Detection decoder of the HM-ViT model on the ego's fused feature: replaces
//   HeteroDecoder.forward (use_upsample=False)   opencood/models/sub_modules/hetero_decoder.py:42-74
//   NaiveDecoder.forward                        opencood/models/sub_modules/naive_decoder.py:63-92
//   camera / lidar cls_head, reg_head (1x1)     hetero_decoder.py:33-40
// i.e. per scene, with the weights of the EGO's modality: 2 * num_layer x (conv3x3 256->256 + BatchNorm(eval) + ReLU), then
// the two 1x1 heads.  BatchNorm is folded into the convolution on the host (hm-vit_b200/decoder.py).
//
//   nchw_to_nhwc_f16_kernel   fused feature fp32 (B,256,H,W) -> fp16 pixel rows [B][H][W][256] (what the TMA boxes want)
//   conv3x3_kernel            TMA-shifted implicit GEMM on tcgen05: one CTA = 16 x 16 output pixels (two M = 128 halves that share
//                             every weight block: the kernel is bound by the L2 -> shared-memory stream, 48 KB per 512 MMA
//                             cycles with one half, 64 KB per 1024 with two) x 256 output channels.
//                             The 3x3 taps are nine K-blocks of the same GEMM: for tap (dy, dx) the A tile is ONE 4-D TMA box
//                             [1 scene][8 rows][16 cols][64 channels] loaded at pixel offset (dy, dx) -- rows outside the
//                             map arrive as zeros (TMA out-of-bounds fill), which is the convolution's zero padding -- and
//                             the B tile the tap's [256 out][64 in] weight block of the ego's modality.  36 K-blocks of 64
//                             through a 3-stage 64 KB ring, two 128 x 256 fp32 accumulators (all 512 columns of tensor
//                             memory), epilogue + bias, ReLU, fp16 pixel rows.  fp16 operands (11-bit significand, like the FFN of the fusion block): bf16
//                             would put ~3e-3 per layer on the logits, the stated tolerance is 1e-3.
//                             The LAST layer (anchor_number == 2, the shipped yaml) keeps its rectified output rows in registers
//                             (fp32) and applies the two 1x1 heads in its epilogue: no fp16 round trip, no heads launch.
//   det_heads_kernel          the 1x1 classification / regression heads on the last feature as their own launch (other anchor
//                             numbers): fp32, thread == pixel.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace hmvit {

struct DecCfg {
  static constexpr int TH = 16, TW = 16;                // output tile: 2 halves of 8 rows x 16 columns = 2 x 128 pixels
  static constexpr int BOX_H = 8;                       // rows of one TMA activation box (one M = 128 half)
  static constexpr int STAGES = 3;
  static constexpr int A_BYTES = 128 * 128;             // one half: 128 pixels x 64 channels fp16
  static constexpr int B_BYTES = 256 * 128;             // 256 output channels x 64 input channels fp16
  static constexpr int STAGE_BYTES = 2 * A_BYTES + B_BYTES; // 64 KB
  static constexpr int KBLOCKS = 9 * 4;                 // 9 taps x 4 channel blocks of 64
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int THREADS = 192;                   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
  static constexpr uint32_t TM_COLS = 512;              // two 128 x 256 fp32 accumulators
};

struct ConvParams {
  int B, H, W;
  const int* ego_mode;         // [B] 0 = camera, 1 = lidar: which weight set a scene uses
  const float* bias;           // [2][256] BatchNorm-folded bias of this layer
  __half* out;                 // [B][H][W][256]
  // last layer with the 1x1 heads in its epilogue (kHeadOut > 0): the layer's output is not stored
  const float* head_w;         // [2][kHeadOut][256]
  const float* head_b;         // [2][kHeadOut]
  float* psm;                  // (B, n_cls, H W)
  float* rm;                   // (B, kHeadOut - n_cls, H W)
  int n_cls;
};

// 4-D tiled TMA load (channels, column, row, scene), completes on an mbarrier; out-of-range elements are zero-filled
HMVIT_DEVINL void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// grid (tiles_x * tiles_y, B).  kHeadOut > 0: the LAST layer -- its rectified output row stays in registers (fp32, not rounded
// to fp16) and feeds the 1x1 classification / regression heads (kHeadOut = 8 x anchor_number outputs, fp32 FMA, weights staged
// in the shared memory the operand ring no longer needs); nothing but psm / rm is written.
template <int kHeadOut>
__global__ void __launch_bounds__(DecCfg::THREADS, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvParams p) {
  using Cfg = DecCfg;
  const int tiles_x = (p.W + Cfg::TW - 1) / Cfg::TW;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int b = blockIdx.y;
  const int h0 = ty * Cfg::TH, w0 = tx * Cfg::TW;
  const int type = p.ego_mode[b] != 0 ? 1 : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* full = bars;                    // [STAGES] TMA -> MMA
  uint64_t* empty = bars + Cfg::STAGES;     // [STAGES] MMA (commit) -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TM_COLS>(tmem_slot);
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_x); tma_prefetch_desc(&tmap_w); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    for (int kb = 0; kb < Cfg::KBLOCKS; ++kb) {
      const int s = kb % Cfg::STAGES;
      mbar_wait(&empty[s], ((kb / Cfg::STAGES) & 1) ^ 1);
      if (elect_one()) {
        const int tap = kb >> 2, cb = kb & 3;
        const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
        tma_load_4d(sa, &tmap_x, &full[s], cb * 64, w0 + dx, h0 + dy, b);
        tma_load_4d(sa + Cfg::A_BYTES, &tmap_x, &full[s], cb * 64, w0 + dx, h0 + 8 + dy, b);
        tma_load_2d(sa + 2 * Cfg::A_BYTES, &tmap_w, &full[s], cb * 64, (type * 9 + tap) * 256);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = umma_idesc(0u, 128, 256);            // fp16 x fp16 -> fp32, M128 N256
    for (int kb = 0; kb < Cfg::KBLOCKS; ++kb) {
      const int s = kb % Cfg::STAGES;
      mbar_wait(&full[s], (kb / Cfg::STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES), sb = sa + 2 * Cfg::A_BYTES;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_ss<2>(tm + hf * 256, umma_desc_sw128(sa + hf * Cfg::A_BYTES + ks * 32), umma_desc_sw128(sb + ks * 32), idesc,
                       (kb | ks) != 0 ? 1u : 0u);
        umma_commit(&empty[s]);
        if (kb == Cfg::KBLOCKS - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------ epilogue: + bias, ReLU, fp16 pixel rows ------------------------------
    const int q = warp & 3;                                         // TMEM lane quadrant this warp may read
    const int pix = q * 32 + lane;                                  // pixel of a half tile == accumulator row
    const float* bias = p.bias + type * 256;
    mbar_wait_sleepy(acc_full, 0);
    tc_fence_after();
    if constexpr (kHeadOut > 0) {
      // every MMA has retired, so every operand stage has been read: stage 0 now holds the head weights of the ego's type
      float* sHw = reinterpret_cast<float*>(smem);                 // [kHeadOut][256]
      const float4* src = reinterpret_cast<const float4*>(p.head_w + static_cast<size_t>(type) * kHeadOut * 256);
      for (int e = threadIdx.x - 64; e < kHeadOut * 64; e += 128) reinterpret_cast<float4*>(sHw)[e] = __ldg(src + e);
      named_bar_sync(1, 128);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {
        const int h = h0 + hf * 8 + (pix >> 4), w = w0 + (pix & 15);
        const bool inside = h < p.H && w < p.W;
        const uint32_t taddr = tm + hf * 256 + (static_cast<uint32_t>(q * 32) << 16);
        float acc[kHeadOut];
#pragma unroll
        for (int o = 0; o < kHeadOut; ++o) acc[o] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c * 32 + k));
            v[k] = __float_as_uint(fmaxf(__uint_as_float(v[k]) + bb.x, 0.f));
            v[k + 1] = __float_as_uint(fmaxf(__uint_as_float(v[k + 1]) + bb.y, 0.f));
          }
#pragma unroll
          for (int o = 0; o < kHeadOut; ++o) {
            const float* wrow = sHw + o * 256 + c * 32;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 ww = *reinterpret_cast<const float4*>(wrow + k4 * 4);       // broadcast read
              acc[o] = fmaf(__uint_as_float(v[k4 * 4]), ww.x, fmaf(__uint_as_float(v[k4 * 4 + 1]), ww.y,
                       fmaf(__uint_as_float(v[k4 * 4 + 2]), ww.z, fmaf(__uint_as_float(v[k4 * 4 + 3]), ww.w, acc[o]))));
            }
          }
        }
        if (inside) {
          const size_t N = static_cast<size_t>(p.H) * p.W, n = static_cast<size_t>(h) * p.W + w;
          const float* hb = p.head_b + type * kHeadOut;
          const int n_reg = kHeadOut - p.n_cls;
#pragma unroll
          for (int o = 0; o < kHeadOut; ++o) {
            const float y = acc[o] + __ldg(hb + o);
            if (o < p.n_cls) p.psm[(static_cast<size_t>(b) * p.n_cls + o) * N + n] = y;
            else p.rm[(static_cast<size_t>(b) * n_reg + (o - p.n_cls)) * N + n] = y;
          }
        }
      }
    } else {
#pragma unroll 1
    for (int c16 = 0; c16 < 16; ++c16) {
      const int hf = c16 >> 3, c = c16 & 7;
      const int h = h0 + hf * 8 + (pix >> 4), w = w0 + (pix & 15);
      const bool inside = h < p.H && w < p.W;
      __half* orow = p.out + ((static_cast<size_t>(b) * p.H + (inside ? h : 0)) * p.W + (inside ? w : 0)) * 256;
      const uint32_t taddr = tm + hf * 256 + (static_cast<uint32_t>(q * 32) << 16);
      uint32_t v[32];
      tmem_ld32(taddr + c * 32, v);
      tmem_ld_wait();
      if (inside) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int ch = c * 32 + u * 8 + k * 2;
            const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + ch));
            const float y0 = fmaxf(__uint_as_float(v[u * 8 + k * 2]) + bb.x, 0.f);
            const float y1 = fmaxf(__uint_as_float(v[u * 8 + k * 2 + 1]) + bb.y, 0.f);
            o[k] = pack_f16x2(y0, y1);
          }
          *reinterpret_cast<uint4*>(orow + c * 32 + u * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TM_COLS>(tm);
  }
}

// fp32 (B, 256, N) -> fp16 [B][N][256]; grid (ceil(N / 64), B), 256 threads
__global__ void __launch_bounds__(256) nchw_to_nhwc_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, int N) {
  __shared__ __half tile[64][256 + 8];
  const int b = blockIdx.y, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, tyy = threadIdx.x >> 6;          // 64 pixels x 4 channel lanes
  const float* src = in + static_cast<size_t>(b) * 256 * N;
  for (int c = tyy; c < 256; c += 4)
    tile[tx][c] = __float2half_rn(n0 + tx < N ? src[static_cast<size_t>(c) * N + n0 + tx] : 0.f);
  __syncthreads();
  __half* dst = out + (static_cast<size_t>(b) * N + n0) * 256;
  for (int e = threadIdx.x; e < 64 * 32; e += 256) {                // 16-byte units
    const int r = e >> 5, u = e & 31;
    if (n0 + r < N) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * 256 + u * 8) = *reinterpret_cast<const uint4*>(&tile[r][u * 8]);
  }
}

constexpr int kDecMaxOut = 32;   // 8 * anchor_number outputs of the two heads (anchor_number <= 4)

struct HeadsParams {
  int B, N, n_cls, n_reg;      // pixels per scene; anchor_number, 7 * anchor_number
  const int* ego_mode;
  const __half* x;             // [B][N][256]
  const float* w;              // [2][n_cls + n_reg][256]
  const float* bias;           // [2][n_cls + n_reg]
  float* psm;                  // (B, n_cls, N)
  float* rm;                   // (B, n_reg, N)
};

// grid (ceil(N / 128), B), 128 threads: thread == pixel
__global__ void __launch_bounds__(128) det_heads_kernel(const HeadsParams p) {
  __shared__ float sw[kDecMaxOut * 256];
  __shared__ float sb[kDecMaxOut];
  const int b = blockIdx.y;
  const int type = p.ego_mode[b] != 0 ? 1 : 0;
  const int no = p.n_cls + p.n_reg;
  for (int e = threadIdx.x; e < no * 256; e += 128) sw[e] = __ldg(p.w + static_cast<size_t>(type) * no * 256 + e);
  if (threadIdx.x < no) sb[threadIdx.x] = __ldg(p.bias + type * no + threadIdx.x);
  __syncthreads();
  const int n = blockIdx.x * 128 + threadIdx.x;
  if (n >= p.N) return;
  float acc[kDecMaxOut];
#pragma unroll
  for (int o = 0; o < kDecMaxOut; ++o) acc[o] = 0.f;
  const uint4* row = reinterpret_cast<const uint4*>(p.x + (static_cast<size_t>(b) * p.N + n) * 256);
#pragma unroll 1
  for (int u = 0; u < 32; ++u) {
    const uint4 v = __ldg(row + u);
    const __half2* h2 = reinterpret_cast<const __half2*>(&v);
    float x[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 f = __half22float2(h2[k]); x[2 * k] = f.x; x[2 * k + 1] = f.y; }
#pragma unroll
    for (int o = 0; o < kDecMaxOut; ++o) {
      if (o < no) {
        const float4 w0 = *reinterpret_cast<const float4*>(&sw[o * 256 + u * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&sw[o * 256 + u * 8 + 4]);
        acc[o] = fmaf(x[0], w0.x, fmaf(x[1], w0.y, fmaf(x[2], w0.z, fmaf(x[3], w0.w,
                 fmaf(x[4], w1.x, fmaf(x[5], w1.y, fmaf(x[6], w1.z, fmaf(x[7], w1.w, acc[o]))))))));
      }
    }
  }
#pragma unroll
  for (int o = 0; o < kDecMaxOut; ++o) {
    if (o < no) {
      const float y = acc[o] + sb[o];
      if (o < p.n_cls) p.psm[(static_cast<size_t>(b) * p.n_cls + o) * p.N + n] = y;
      else p.rm[(static_cast<size_t>(b) * p.n_reg + (o - p.n_cls)) * p.N + n] = y;
    }
  }
}

}  // namespace hmvit
