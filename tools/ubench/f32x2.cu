// Micro-benchmark: packed f32x2 add / fma issue rate on sm_100a vs scalar FADD/FFMA.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack(float x, float y) {
  return ((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(x);
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a0, float b0, int iters) {
  unsigned long long acc[8];
  float facc[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = pack(a0 + i + threadIdx.x, a0 - i);
#pragma unroll
  for (int i = 0; i < 16; ++i) facc[i] = a0 + i + threadIdx.x;
  const float bf = b0 + threadIdx.x * 1e-9f;
  const unsigned long long b = pack(bf, bf * 0.5f), c = pack(bf * 0.25f, bf * 0.125f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = add2(acc[i], b);      // 8 packed adds = 16 flops-lanes
      } else if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma2(acc[i], b, c);
      } else if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 16; ++i) facc[i] = facc[i] + bf;       // 16 scalar adds
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) facc[i] = fmaf(facc[i], bf, 0.3f);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
#pragma unroll
  for (int i = 0; i < 16; ++i) s += facc[i];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 4096;
  k<MODE><<<148 * 8, 256>>>(out, 1.0f, 0.999f, iters);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, 1.0f, 0.999f, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double lanes = (double)148 * 8 * 256 * iters * 8 * 16;  // scalar-equivalent f32 operations per thread-loop
  printf("%-28s %.3f ms  %.2f T f32-ops/s  = %.1f f32 ops per clk per SM at 1965 MHz\n", name, ms, lanes / ms / 1e9, lanes / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<2>("FADD scalar");
  run<0>("add.f32x2");
  run<3>("FFMA scalar");
  run<1>("fma.rn.f32x2");
  return 0;
}
