// wmma operand-fragment loads from the shared state space (used by the backward kernels).
#pragma once
#include "common.cuh"
#include <mma.h>
#include <type_traits>

namespace hmvit {

// wmma fragment load from SHARED memory.  nvcuda::wmma::load_matrix_sync takes a generic pointer and compiled (CUDA 12.9,
// sm_100a) to scalar generic LD.E instructions for most call sites -- 107 M of them per launch of the attention backward
// (ncu source view).  The register layout of a fragment does not depend on the state space, so the same fragments feed
// wmma::mma_sync.
template <class Use, class Layout>
HMVIT_DEVINL void wmma_load_shared(nvcuda::wmma::fragment<Use, 16, 16, 16, __nv_bfloat16, Layout>& f, const __nv_bfloat16* p, unsigned ld) {
  using namespace nvcuda::wmma;
  static_assert(sizeof(f.x) == 16, "bf16 m16n16k16 operand fragment: four 32-bit registers");
  uint32_t r0, r1, r2, r3;
  const uint32_t addr = smem_u32(p);
  // The two TRANSPOSED layouts (A col-major, B row-major: the contraction index is the slow one in memory) compile, as
  // wmma.load, to eight 16-bit shared loads + four PRMT per fragment -- a quarter of the attention backward's instructions.
  // ldmatrix.trans delivers the same registers in one instruction: four 8 x 8 tiles, tile i addressed by lanes 8 i .. 8 i + 7,
  // each lane receiving (stored row 2 t, 2 t + 1; stored column g) of tile i in register i (g = lane / 4, t = lane % 4), which
  // is the mma.m16n8k16 operand layout the m16n16k16 fragment is made of: A = (m 0-7 | 8-15) x (k 0-7 | 8-15) in the order
  // a0 (m lo, k lo), a1 (m hi, k lo), a2 (m lo, k hi), a3 (m hi, k hi); B = b0 (k lo, n lo), b1 (k hi, n lo), b2 (k lo, n hi),
  // b3 (k hi, n hi).  Stored rows are the contraction index k; requires 16-byte aligned rows (ld % 8 == 0).
  constexpr bool kTransA = std::is_same<Use, matrix_a>::value && std::is_same<Layout, col_major>::value;
  constexpr bool kTransB = std::is_same<Use, matrix_b>::value && std::is_same<Layout, row_major>::value;
  if constexpr (kTransA || kTransB) {
    const uint32_t lane = threadIdx.x & 31u, tile = lane >> 3, r = lane & 7u;
    const uint32_t kblk = kTransA ? (tile >> 1) : (tile & 1u), mnblk = kTransA ? (tile & 1u) : (tile >> 1);
    const uint32_t a = addr + ((kblk * 8u + r) * ld + mnblk * 8u) * 2u;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
  } else
  if constexpr (std::is_same<Use, matrix_a>::value && std::is_same<Layout, row_major>::value)
    asm volatile("wmma.load.a.sync.aligned.row.m16n16k16.shared.bf16 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr), "r"(ld));
  else if constexpr (std::is_same<Use, matrix_a>::value)
    asm volatile("wmma.load.a.sync.aligned.col.m16n16k16.shared.bf16 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr), "r"(ld));
  else if constexpr (std::is_same<Layout, row_major>::value)
    asm volatile("wmma.load.b.sync.aligned.row.m16n16k16.shared.bf16 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr), "r"(ld));
  else
    asm volatile("wmma.load.b.sync.aligned.col.m16n16k16.shared.bf16 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr), "r"(ld));
  uint32_t* x = reinterpret_cast<uint32_t*>(f.x);
  x[0] = r0; x[1] = r1; x[2] = r2; x[3] = r3;
}

}  // namespace hmvit
