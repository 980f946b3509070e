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
