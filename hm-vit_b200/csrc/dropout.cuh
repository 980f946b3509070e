// Dropout of the training path (the shipped yaml trains with drop_out: 0.1):
//   nn.Dropout after a_linears      opencood/models/sub_modules/hetero_fusion.py:66
//   nn.Dropout x 2 in the FFN       opencood/models/base_transformer.py:186-190
// The mask is a pure function of (seed, stream, element index) -- Philox4x32-10, one 128-bit block per 4 consecutive
// elements of a channel-major row -- so it is never stored: the backward regenerates it, and the tests export it
// (a == NULL) to hand the identical mask to the CPU oracle.  Bit parity with torch's own RNG stream is not attempted.
#pragma once
#include "common.cuh"

namespace hmvit {

HMVIT_DEVINL void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
HMVIT_DEVINL void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

struct DropoutParams {
  const float* a;              // cm fp32 [agents][256][N], or null: a == 1 (mask export)
  const float* resid;          // optional cm fp32 added AFTER the dropout, or null
  float* out;                  // may alias a or resid
  int L, N;
  const int* record_len;       // [B]
  int ego_only;
  uint32_t seed_lo, seed_hi;   // Philox key
  uint32_t stream;             // which dropout site / stage / call (Philox counter word 2)
  uint32_t threshold;          // keep iff random u32 >= threshold (threshold = p * 2^32)
  float scale;                 // 1 / (1 - p)
};

// grid (ceil(N / 4 / 256), 256 channels, agents); thread = 4 consecutive tokens of one channel row
__global__ void __launch_bounds__(256) dropout_cm_kernel(const DropoutParams p) {
  const int a = blockIdx.z, ch = blockIdx.y;
  const int b = a / p.L, l = a - b * p.L;
  if (l >= min(p.record_len[b], p.L) || (p.ego_only && l != 0)) return;
  const int q4 = blockIdx.x * blockDim.x + threadIdx.x;       // group of 4 tokens
  const int tok = q4 * 4;
  if (tok >= p.N) return;
  const size_t row = (static_cast<size_t>(a) * kC + ch) * p.N;
  const unsigned long long idx = (row + tok) >> 2;            // Philox block index (N % 4 == 0)
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), p.stream, 0u, p.seed_lo, p.seed_hi, r);
  float4 v = p.a != nullptr ? *reinterpret_cast<const float4*>(p.a + row + tok) : make_float4(1.f, 1.f, 1.f, 1.f);
  v.x = r[0] >= p.threshold ? v.x * p.scale : 0.f;
  v.y = r[1] >= p.threshold ? v.y * p.scale : 0.f;
  v.z = r[2] >= p.threshold ? v.z * p.scale : 0.f;
  v.w = r[3] >= p.threshold ? v.w * p.scale : 0.f;
  if (p.resid != nullptr) {
    const float4 rr = *reinterpret_cast<const float4*>(p.resid + row + tok);
    v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
  }
  *reinterpret_cast<float4*>(p.out + row + tok) = v;
}

}  // namespace hmvit
