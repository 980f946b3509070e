// PointPillar front end (the LiDAR branch in front of the fusion): replaces
//   PillarVFE.forward            opencood/models/sub_modules/pillar_vfe.py:100-146  (single PFN layer, use_absolute_xyz,
//                                                                                   no distance feature: the shipped yaml)
//   PFNLayer.forward             pillar_vfe.py:32-54   (Linear without bias + BatchNorm1d(eval) + ReLU + max over the points)
//   PointPillarScatter.forward   opencood/models/sub_modules/point_pillar_scatter.py:15-48
// in one pass: pillar_vfe_scatter_kernel, warp == pillar.  The reference materialises the (M, 32, 10) augmented points, the
// (M, 32, 64) linear output, its normalised / rectified copies and the (M, 64) maxima, then scatters per agent in a Python
// loop; here a pillar's 32 points are read once (512 B), its 10-feature rows live in shared memory (1.25 KB per warp), every
// lane owns two of the 64 output channels (weights in registers: 2 x 10, BatchNorm folded on the host), and the 64 maxima
// go straight into the dense canvas -- (n, 64, ny, nx), or channels-last [n][ny][nx][64] (what cuDNN's tensor-core
// convolutions want: one coalesced 256-byte row per pillar here, no NCHW -> NHWC pass there).  HBM-bound by construction: 512 B + 16 B read and 256 B written per
// pillar, plus the zero fill of the canvas (done by the C entry with cudaMemsetAsync).
//
// Padded point slots: the reference zeroes the FEATURES of the slots >= num_points (pillar_vfe.py:139-141), not their
// output, so such a slot still contributes relu(BN(0)) = relu(b') to the maximum -- reproduced.
#pragma once
#include "common.cuh"

namespace hmvit {

constexpr int kPfnIn = 10, kPfnOut = 64, kPfnMaxPts = 32;

struct PillarParams {
  int M, P;                    // pillars, point slots per pillar (<= 32)
  const float* pts;            // [M][P][4] x, y, z, intensity (padded slots arbitrary)
  const int* coords;           // [M][4] agent, z, y, x
  const int* npts;             // [M]
  const float* w;              // [64][10] BatchNorm-folded weight
  const float* b;              // [64] BatchNorm-folded shift
  float vx, vy, vz, ox, oy, oz;   // voxel size and centre offset (size / 2 + range_min)
  int nx, ny, n_agents;
  float* canvas;               // (n_agents, 64, ny, nx) or, channels_last, [n_agents][ny][nx][64]; zero-filled by the caller
  int channels_last;
};

// grid: ceil(M / 8) blocks of 256 threads (8 warps == 8 pillars)
__global__ void __launch_bounds__(256) pillar_vfe_scatter_kernel(const PillarParams p) {
  __shared__ __align__(16) float sF[8][kPfnMaxPts][kPfnIn + 2];          // per warp: the augmented points (row padded to 12 floats = 3 x 16 B)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * 8 + warp;
  if (m >= p.M) return;
  const int n = min(max(p.npts[m], 0), p.P);
  const int4 c = *reinterpret_cast<const int4*>(p.coords + static_cast<size_t>(m) * 4);
  // this lane's two output channels
  float w0[kPfnIn], w1[kPfnIn];
#pragma unroll
  for (int k = 0; k < kPfnIn; ++k) { w0[k] = __ldg(p.w + (2 * lane) * kPfnIn + k); w1[k] = __ldg(p.w + (2 * lane + 1) * kPfnIn + k); }
  const float b0 = __ldg(p.b + 2 * lane), b1 = __ldg(p.b + 2 * lane + 1);
  // lane == point slot: load, mean of the xyz of ALL P slots' stored values divided by num_points like the reference
  // (voxel_features[:, :, :3].sum(dim=1) / num_points: padded slots are zeros in the reference's input)
  float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < p.P) pt = __ldg(reinterpret_cast<const float4*>(p.pts) + static_cast<size_t>(m) * p.P + lane);
  float sx = pt.x, sy = pt.y, sz = pt.z;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  const float inv = 1.0f / static_cast<float>(p.npts[m]);
  const float mx = sx * inv, my = sy * inv, mz = sz * inv;
  const float cx = static_cast<float>(c.w) * p.vx + p.ox, cy = static_cast<float>(c.z) * p.vy + p.oy, cz = static_cast<float>(c.y) * p.vz + p.oz;
  const bool live = lane < n;
  float* f = sF[warp][lane];
  f[0] = live ? pt.x : 0.f; f[1] = live ? pt.y : 0.f; f[2] = live ? pt.z : 0.f; f[3] = live ? pt.w : 0.f;
  f[4] = live ? pt.x - mx : 0.f; f[5] = live ? pt.y - my : 0.f; f[6] = live ? pt.z - mz : 0.f;
  f[7] = live ? pt.x - cx : 0.f; f[8] = live ? pt.y - cy : 0.f; f[9] = live ? pt.z - cz : 0.f;
  __syncwarp();
  // padded slots (n < P) contribute relu(b'); live slots their own value
  float best0 = n < p.P ? fmaxf(b0, 0.f) : 0.f, best1 = n < p.P ? fmaxf(b1, 0.f) : 0.f;
  for (int q = 0; q < n; ++q) {
    const float4 fa = *reinterpret_cast<const float4*>(&sF[warp][q][0]);
    const float4 fb = *reinterpret_cast<const float4*>(&sF[warp][q][4]);
    const float2 fc = *reinterpret_cast<const float2*>(&sF[warp][q][8]);
    const float x[kPfnIn] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w, fc.x, fc.y};
    // (the reference's nn.Linear accumulates the 10 products in its own order; fp32 either way)
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int k = 0; k < kPfnIn; ++k) { y0 = fmaf(x[k], w0[k], y0); y1 = fmaf(x[k], w1[k], y1); }
    best0 = fmaxf(best0, y0 + b0);                         // relu folded into the maximum (best starts at >= 0 ...)
    best1 = fmaxf(best1, y1 + b1);
  }
  // ... except for a full pillar (n == P), whose maximum may be negative before the ReLU: clamp
  best0 = fmaxf(best0, 0.f); best1 = fmaxf(best1, 0.f);
  const int a = c.x;
  if (a < 0 || a >= p.n_agents || c.y != 0 || c.z < 0 || c.z >= p.ny || c.w < 0 || c.w >= p.nx) return;    // malformed coordinate: dropped
  // the reference's flat cell index: z + y * nx + x (point_pillar_scatter.py:32-34), nz == 1
  const size_t cell = static_cast<size_t>(c.y) + static_cast<size_t>(c.z) * p.nx + c.w;
  if (p.channels_last) {                                  // [n][ny][nx][64]: one coalesced 256-byte row per pillar
    float* dst = p.canvas + (static_cast<size_t>(a) * p.ny * p.nx + cell) * kPfnOut + 2 * lane;
    *reinterpret_cast<float2*>(dst) = make_float2(best0, best1);
  } else {                                                // (n, 64, ny, nx)
    float* dst = p.canvas + (static_cast<size_t>(a) * kPfnOut + 2 * lane) * (static_cast<size_t>(p.ny) * p.nx) + cell;
    dst[0] = best0;
    dst[static_cast<size_t>(p.ny) * p.nx] = best1;
  }
}

}  // namespace hmvit
