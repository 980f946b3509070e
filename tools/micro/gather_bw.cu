// Micro-benchmark: achievable bandwidth of 256-byte row gathers (16 lanes x 16 B per row) on B200,
// as a function of loads in flight per thread, resident warps and working-set size (L2 vs HBM),
// and the same rows fetched with cp.async.bulk (one 256-byte bulk copy per lane) into shared memory.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int U>
__global__ void gather_ldg(const uint4* __restrict__ buf, uint32_t nrows, int iters, uint4* out) {
  const uint32_t hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, u16 = threadIdx.x & 15;
  uint4 acc = make_uint4(0, 0, 0, 0);
  uint32_t seed = hw * 2654435761u;
  for (int it = 0; it < iters; ++it) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      seed = hash32(seed + k + 1);
      v[k] = __ldg(buf + static_cast<size_t>(seed % nrows) * 16 + u16);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) { acc.x ^= v[k].x; acc.y += v[k].y; acc.z ^= v[k].z; acc.w += v[k].w; }
  }
  if (acc.x == 0x12345678u) out[0] = acc;
}

// one 256-byte bulk copy per lane and iteration, R rounds in flight (ring of R smem slots per thread)
template <int R>
__global__ void gather_bulk(const uint4* __restrict__ buf, uint32_t nrows, int iters, uint4* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar[R];
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) { for (int r = 0; r < R; ++r) { uint32_t a = (uint32_t)__cvta_generic_to_shared(&bar[r]); asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(nt)); } asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  uint32_t seed = (blockIdx.x * nt + tid) * 2654435761u;
  uint32_t acc = 0;
  for (int it = 0; it < iters + R; ++it) {
    const int r = it % R;
    const uint32_t ba = (uint32_t)__cvta_generic_to_shared(&bar[r]);
    uint8_t* slot = smem + (static_cast<size_t>(r) * nt + tid) * 256;
    if (it >= R) {
      const uint32_t par = ((it / R) - 1) & 1;
      uint32_t ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0,1,0,p;\n\t}" : "=r"(ok) : "r"(ba), "r"(par) : "memory");
      acc += *reinterpret_cast<uint32_t*>(slot);
    }
    if (it < iters) {
      seed = hash32(seed + 1);
      const uint4* src = buf + static_cast<size_t>(seed % nrows) * 16;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(256) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"((uint32_t)__cvta_generic_to_shared(slot)), "l"(src), "r"(256), "r"(ba) : "memory");
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ba) : "memory");
    }
  }
  if (acc == 0x12345678u) out[0] = make_uint4(acc, 0, 0, 0);
}

template <int U>
void run_ldg(const uint4* buf, uint32_t nrows, int ctas_per_sm, int threads, uint4* out, const char* tag) {
  const int iters = 2048 / U;
  const int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather_ldg<U><<<grid, threads>>>(buf, nrows, iters, out);
  cudaEventRecord(e0);
  gather_ldg<U><<<grid, threads>>>(buf, nrows, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = double(grid) * threads * iters * U * 16;
  printf("%s ldg U=%2d ctas/sm=%d threads=%d (warps/sm=%d, %d KB in flight/SM): %.0f GB/s\n", tag, U, ctas_per_sm, threads,
         ctas_per_sm * threads / 32, ctas_per_sm * threads * U * 16 / 1024, bytes / ms / 1e6);
}
template <int R>
void run_bulk(const uint4* buf, uint32_t nrows, int ctas_per_sm, int threads, uint4* out, const char* tag) {
  const int iters = 512;
  const int grid = 148 * ctas_per_sm;
  const size_t smem = size_t(R) * threads * 256;
  cudaFuncSetAttribute(gather_bulk<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  gather_bulk<R><<<grid, threads, smem>>>(buf, nrows, iters, out);
  cudaEventRecord(e0);
  gather_bulk<R><<<grid, threads, smem>>>(buf, nrows, iters, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = double(grid) * threads * iters * 256;
  printf("%s bulk256 R=%d ctas/sm=%d threads=%d (%zu KB in flight/SM): %.0f GB/s  [%s]\n", tag, R, ctas_per_sm, threads,
         ctas_per_sm * smem / 1024, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  uint4 *big, *out;
  const size_t big_bytes = 700ull << 20;
  cudaMalloc(&big, big_bytes); cudaMalloc(&out, 64);
  cudaMemset(big, 1, big_bytes);
  for (int pass = 0; pass < 2; ++pass) {
    const uint32_t nrows = pass == 0 ? (40u << 20) / 256 : (uint32_t)(big_bytes / 256);
    const char* tag = pass == 0 ? "[40MB  L2 ]" : "[700MB HBM]";
    run_ldg<4>(big, nrows, 2, 128, out, tag);
    run_ldg<8>(big, nrows, 2, 128, out, tag);
    run_ldg<16>(big, nrows, 2, 128, out, tag);
    run_ldg<16>(big, nrows, 4, 128, out, tag);
    run_ldg<16>(big, nrows, 4, 256, out, tag);
    run_ldg<32>(big, nrows, 2, 256, out, tag);
    run_ldg<8>(big, nrows, 8, 256, out, tag);
    run_bulk<2>(big, nrows, 2, 128, out, tag);
    run_bulk<4>(big, nrows, 2, 64, out, tag);
    run_bulk<2>(big, nrows, 2, 256, out, tag);
    run_bulk<3>(big, nrows, 1, 256, out, tag);
  }
  return 0;
}
