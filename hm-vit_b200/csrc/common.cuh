// Shared sm_100a primitives for the HM-ViT fusion kernels: mbarrier, TMA, tcgen05 (TMEM alloc,
// UMMA issue/commit, TMEM load/store), swizzle-128B address math, small numeric helpers.
// Everything here is raw inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef HMVIT_DEVINL
#define HMVIT_DEVINL __device__ __forceinline__
#endif

namespace hmvit {

constexpr int kC = 256;        // feature channels (input_dim == mlp_dim == 256)
constexpr int kHeads = 8;      // heads = kC / 32
constexpr int kDh = 32;        // dim_head
constexpr int kWin = 8;        // window size
constexpr int kS = 64;         // tokens per group (window^2)

// ------------------------------------------------------------------------------------------
// generic helpers
// ------------------------------------------------------------------------------------------
HMVIT_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 1024-byte aligned start of the dynamic shared memory.  The pad is computed on the 32-bit shared-window
// address and ADDED to the __shared__ array, so the compiler keeps the pointer in the shared address space
// (LDS / STS).  Rounding a uintptr_t and casting back yields a generic pointer: every shared access then
// compiles to a generic LD / ST through the global-memory path (seen in SASS as LD.E / ST.E + long scoreboard).
HMVIT_DEVINL uint8_t* smem_align1024(uint8_t* smem_raw) {
  const uint32_t base = smem_u32(smem_raw);
  return smem_raw + (((base + 1023u) & ~1023u) - base);
}

HMVIT_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}
// two fp32 -> packed fp16 (lo in the low half), round to nearest, saturating at +-65504 instead of overflowing to inf
HMVIT_DEVINL uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
HMVIT_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
HMVIT_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// round-to-nearest (ties away) fp32 -> tf32, result kept in an fp32 container
HMVIT_DEVINL float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

HMVIT_DEVINL float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// GELU (erf form) in 9 instructions with ONE MUFU: with u = |x|,
//   gelu(x) = x (1 + erf(x / sqrt 2)) / 2 = (x + u) / 2 - (u / 2) erfc(u / sqrt 2),   erfc(u / sqrt 2) = 2^(-u q(u)),
// q a degree-4 polynomial fitted to -log2(erfc(u / sqrt 2)) / u on [0, 6.5] (weighted by the slope of the result; beyond 6.5
// the exponential underflows to 0, which is the limit).  |abs err| <= 9.3e-7 over [-12, 12] evaluated in fp32 -- 500 x below
// the fp16 rounding applied to the result -- and the negative tail keeps its relative accuracy (no 1 - erf cancellation).
// The Abramowitz-Stegun 7.1.26 form used before needed two MUFU (rcp, ex2) and ~15 instructions; the GELU feed of the chain
// kernel is MUFU- and issue-bound (128 x 256 activations per tile).
HMVIT_DEVINL float gelu_erf_fast(float x) {
  const float u = fabsf(x);
  float q = fmaf(u, 4.884928348474205e-4f, -7.197308354079723e-3f);
  q = fmaf(u, q, 5.213385447859764e-2f);
  q = fmaf(u, q, 4.596169888973236e-1f);
  q = fmaf(u, q, 1.1509909629821777f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * q));
  const float h = 0.5f * u;
  return fmaf(-h, e, fmaf(0.5f, x, h));
}

// ------------------------------------------------------------------------------------------
// 128-byte swizzle (UMMA canonical K-major SWIZZLE_128B layout == what TMA SWIZZLE_128B writes)
// A [rows x 128 B] chunk: row r at r*128, its eight 16-byte units XORed with (r & 7).
// The chunk base must be 1024-byte aligned.
// ------------------------------------------------------------------------------------------
HMVIT_DEVINL uint32_t sw128_offset(uint32_t row, uint32_t unit16 /*0..7*/) {
  return row * 128u + ((unit16 ^ (row & 7u)) << 4);
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
HMVIT_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
HMVIT_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
HMVIT_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
HMVIT_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
HMVIT_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a broken pipeline traps (-> launch error surfaced through the C-ABI) instead of
// hanging the GPU box.
HMVIT_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { asm volatile("trap;"); }
  }
}

// wait of a consumer that usually arrives early (softmax warps waiting for S, the MMA warp waiting for P): the try_wait
// carries a suspend-time hint, so the warp sleeps in hardware instead of spinning through the issue slots of the
// producer warps it is waiting for (a plain spin loop was 20 % of all executed instructions of the fused attention)
HMVIT_DEVINL void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(200000u)
        : "memory");
    if (++spins > (1u << 15)) { asm volatile("trap;"); }   // (each try may sleep up to the hint: ~seconds in total)
  } while (!ok);
}

// non-blocking phase test
HMVIT_DEVINL bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait of a role that is idle for a long time (an item's duration): sleeps between polls so that its spin loop does not
// take issue slots from the working warps
HMVIT_DEVINL void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(128);
    if (++spins > (1u << 24)) { asm volatile("trap;"); }
  }
}

// one lane of a CONVERGED warp (elect.sync): the warp keeps executing uniformly around a single-thread instruction,
// so the compiler can hold descriptors / addresses in uniform registers (an `if (lane == 0)` region makes every
// operand of a tcgen05.mma go through a per-instruction R2UR loop: ~100 issue cycles per MMA, measured)
HMVIT_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// volatile shared-memory accesses: keep a hand-scheduled batch of loads / stores in program order
HMVIT_DEVINL float4 lds_f4(const void* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}
HMVIT_DEVINL uint4 lds_u4(const void* p) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)));
  return v;
}
HMVIT_DEVINL uint4 lds_u4_addr(uint32_t addr) {       // same, from a 32-bit shared-window address
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
HMVIT_DEVINL void sts_u4_addr(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
HMVIT_DEVINL void sts_u4(void* p, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(p)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// make generic-proxy smem writes visible to the async proxy (TMA / UMMA operand reads)
HMVIT_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// TMA (2-D tiled load, completes on an mbarrier)
// ------------------------------------------------------------------------------------------
HMVIT_DEVINL void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
HMVIT_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA 3-D tiled store (shared -> global, bulk async group of the issuing thread); out-of-range box rows are clipped
HMVIT_DEVINL void tma_store_3d(const void* tmap, const void* smem_src, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
HMVIT_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kN> HMVIT_DEVINL void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kN) : "memory"); }
template <int kN> HMVIT_DEVINL void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(kN) : "memory"); }

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA, commit, TMEM <-> registers
// ------------------------------------------------------------------------------------------
template <uint32_t kCols>
HMVIT_DEVINL void tmem_alloc(uint32_t* smem_dst) {   // whole warp, .sync.aligned
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM columns: power of 2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
HMVIT_DEVINL void tmem_dealloc(uint32_t taddr) {     // whole warp (the allocating one)
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
HMVIT_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
HMVIT_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, 8-row groups 1024 B apart.
// (bit layout: cute::UMMA::SmemDescriptor -- start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
//  version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2)
HMVIT_DEVINL uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;             // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO
  d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A/B K-major.
// fmt: 0 = F16, 1 = BF16 (kind::f16), 2 = TF32 (kind::tf32)
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; single elected thread issues.
template <int kES>
HMVIT_DEVINL void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kES == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem]^T, bf16 (A: one row per TMEM lane, two K elements per 32-bit column, 8 columns
// per K = 16 instruction).  The A operand never crosses the shared-memory port.
HMVIT_DEVINL void umma_ts_bf16(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued UMMAs of this thread have completed
HMVIT_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (one row per thread)
HMVIT_DEVINL void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 16-column variants (one row per thread, 16 consecutive fp32 columns)
HMVIT_DEVINL void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
HMVIT_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
HMVIT_DEVINL void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
HMVIT_DEVINL void tmem_ld1(uint32_t taddr, uint32_t (&r)[1]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
HMVIT_DEVINL void tmem_st1(uint32_t taddr, const uint32_t (&r)[1]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(r[0]) : "memory");
}
HMVIT_DEVINL void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
HMVIT_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
HMVIT_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
HMVIT_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// warp geometry shared by the attention gather and the stand-alone warp / ROI-mask kernels
// ------------------------------------------------------------------------------------------
// Source-pixel map of one (source j -> target i) pair, see DESIGN.md "warp":
//   src = Ainv * ((u, v) - c - t) + c,   c = (W/2, H/2),   t = T[:2,3] / (voxel * downsample)
// evaluated in fp64 from the 4x4 fp32 pose (rows {0,1} x cols {0,1,3}), like the oracle.
struct WarpMap {
  double a00, a01, a10, a11;   // Ainv
  double ox, oy;               // c + t
  double cx, cy;               // c
};
HMVIT_DEVINL WarpMap make_warp_map(const float* __restrict__ T44, int H, int W, double cell) {
  WarpMap m;
  const double A00 = T44[0], A01 = T44[1], A10 = T44[4], A11 = T44[5];
  const double tx = static_cast<double>(T44[3]) / cell, ty = static_cast<double>(T44[7]) / cell;
  const double det = A00 * A11 - A01 * A10;
  const double inv = 1.0 / det;
  m.a00 = A11 * inv;  m.a01 = -A01 * inv;
  m.a10 = -A10 * inv; m.a11 = A00 * inv;
  m.cx = 0.5 * W; m.cy = 0.5 * H;
  m.ox = m.cx + tx; m.oy = m.cy + ty;
  return m;
}
HMVIT_DEVINL void warp_src(const WarpMap& m, int u, int v, double& sx, double& sy) {
  const double du = static_cast<double>(u) - m.ox, dv = static_cast<double>(v) - m.oy;
  sx = m.a00 * du + m.a01 * dv + m.cx;
  sy = m.a10 * du + m.a11 * dv + m.cy;
}
// nearest-neighbour visibility (grid_sample mode='nearest', zeros padding): rint == half-to-even
HMVIT_DEVINL bool warp_visible(double sx, double sy, int H, int W) {
  const double rx = rint(sx), ry = rint(sy);
  return rx >= 0.0 && rx <= static_cast<double>(W - 1) && ry >= 0.0 && ry <= static_cast<double>(H - 1);
}
// bilinear taps: linear index of the (y0,x0) corner and 4 weights (zero for out-of-range taps)
struct Taps {
  int x0, y0;
  float w00, w01, w10, w11;    // (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1)
};
HMVIT_DEVINL Taps make_taps(double sx, double sy, int H, int W) {
  Taps t;
  const double fx = floor(sx), fy = floor(sy);
  // clamp far-away coordinates so the int conversion is defined; such taps get zero weight anyway
  const double cxl = fmin(fmax(fx, -2.0), static_cast<double>(W)), cyl = fmin(fmax(fy, -2.0), static_cast<double>(H));
  t.x0 = static_cast<int>(cxl); t.y0 = static_cast<int>(cyl);
  const double wx1 = sx - fx, wy1 = sy - fy, wx0 = 1.0 - wx1, wy0 = 1.0 - wy1;
  const bool inx0 = fx >= 0.0 && fx <= static_cast<double>(W - 1), inx1 = fx + 1.0 >= 0.0 && fx + 1.0 <= static_cast<double>(W - 1);
  const bool iny0 = fy >= 0.0 && fy <= static_cast<double>(H - 1), iny1 = fy + 1.0 >= 0.0 && fy + 1.0 <= static_cast<double>(H - 1);
  t.w00 = (inx0 && iny0) ? static_cast<float>(wx0 * wy0) : 0.0f;
  t.w01 = (inx1 && iny0) ? static_cast<float>(wx1 * wy0) : 0.0f;
  t.w10 = (inx0 && iny1) ? static_cast<float>(wx0 * wy1) : 0.0f;
  t.w11 = (inx1 && iny1) ? static_cast<float>(wx1 * wy1) : 0.0f;
  return t;
}

}  // namespace hmvit
