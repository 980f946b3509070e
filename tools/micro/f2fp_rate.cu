// Issue rate of cvt.rn.bf16x2.f32 (F2FP.BF16.F32.PACK_AB) against FADD / integer rounding, per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f2fp_rate f2fp_rate.cu && ./f2fp_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.37f + i;
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      if (MODE == 0) {
        uint32_t r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i + 1]), "f"(a[i]));
        acc ^= r;
      } else if (MODE == 1) {
        float r;
        asm volatile("add.f32 %0, %1, %2;" : "=f"(r) : "f"(a[i + 1]), "f"(a[i]));
        acc ^= __float_as_uint(r);
      } else {
        uint32_t u0 = __float_as_uint(a[i]), u1 = __float_as_uint(a[i + 1]);
        u0 += 0x7fffu + ((u0 >> 16) & 1u); u1 += 0x7fffu + ((u1 >> 16) & 1u);
        acc ^= __byte_perm(u0, u1, 0x7632);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] += 1.0f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps) {
  uint32_t* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  k<MODE><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-10s warps/SM %2d: %.2f cycles per loop body (8 ops + 16 FADD) per warp-iteration\n", name, warps, double(c) / iters);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4, 8, 16}) { run<0>("f2fp", w); run<1>("fadd", w); run<2>("int-rne", w); }
  return 0;
}
