// Bring-up probe for UMMA operand layouts used by the tcgen05 attention kernel (diagnostics only):
//   D[128 x N] (fp32) = A[128 x 64] (bf16, K-major SW128) * B
// with B either K-major ([N][64], 128-byte rows) or MN-major ([64 k][N <= 64], 128-byte rows, the
// layout in which gathered V tiles are written: one row per key, channels contiguous).
#pragma once
#include "common.cuh"

namespace hmvit {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                  int N, int b_mn_major /* bit 0: B MN-major, bit 1: A operand from TMEM */, uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_align1024(smem_raw);
  uint8_t* sA = smem;              // 128 rows x 128 B
  uint8_t* sB = smem + 16384;      // up to 128 rows x 128 B (K-major) / 2 MN blocks of 64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A: row r = tid, 8 units of 16 B
  for (int u = 0; u < 8; ++u)
    *reinterpret_cast<uint4*>(sA + sw128_offset(tid, u)) = *reinterpret_cast<const uint4*>(A + tid * 64 + u * 8);
  // B: rows of 128 B (K-major: row = n, N rows; MN-major: row = k, 64 rows, N*2 bytes used, rest zero)
  const int brows = (b_mn_major & 1) ? 128 : N;    // MN-major: 2 blocks of 64 k-rows (block = 64 n)
  for (int e = tid; e < brows * 8; e += 128) {
    const int r = e >> 3, u = e & 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(b_mn_major & 1)) v = *reinterpret_cast<const uint4*>(Bm + r * 64 + u * 8);
    else {
      const int blk = r >> 6, k = r & 63, n0 = blk * 64 + u * 8;
      if (n0 < N) v = *reinterpret_cast<const uint4*>(Bm + k * N + n0);
    }
    *reinterpret_cast<uint4*>(sB + (r >> 6) * ((b_mn_major & 1) ? 8192 : 0) + sw128_offset((b_mn_major & 1) ? (r & 63) : r, u)) = v;
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<256>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const int a_tmem = (b_mn_major >> 1) & 1;
  b_mn_major &= 1;
  if (a_tmem) {
    // A row of this thread (64 bf16 = 32 packed words) into TMEM columns [64, 96)
    uint32_t w[32];
    for (int k = 0; k < 32; ++k) w[k] = reinterpret_cast<const uint32_t*>(A + tid * 64)[k];
    tmem_st32(tm + (static_cast<uint32_t>(warp * 32) << 16) + 128, w);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = umma_idesc(1u, 128, N) | (b_mn_major ? (1u << 16) : 0u);
    for (int ks = 0; ks < 4; ++ks) {
      uint64_t ad = umma_desc_sw128(smem_u32(sA) + ks * 32);
      uint64_t bd;
      if (!b_mn_major) bd = umma_desc_sw128(smem_u32(sB) + ks * 32);
      else {
        bd = 0;
        bd |= static_cast<uint64_t>(((smem_u32(sB) + ks * kstep_bytes) & 0x3FFFFu) >> 4);
        bd |= static_cast<uint64_t>(lbo >> 4) << 16;
        bd |= static_cast<uint64_t>(sbo >> 4) << 32;
        bd |= static_cast<uint64_t>(1) << 46;
        bd |= static_cast<uint64_t>(2) << 61;
      }
      if (a_tmem) {
        const uint32_t at = tm + 128 + ks * 8;
        const uint32_t acc = ks != 0 ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
            ::"r"(tm), "r"(at), "l"(bd), "r"(idesc), "r"(acc)
            : "memory");
      } else
      umma_ss<2>(tm, ad, bd, idesc, ks != 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int k = 0; k < 16; ++k) D[tid * N + c0 + k] = __uint_as_float(r[k]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tm); }
}

}  // namespace hmvit
