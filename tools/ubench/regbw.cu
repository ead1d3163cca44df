// Micro-benchmark: are packed FP32 instructions limited by register-file bandwidth?  Each variant runs 8 independent chains per
// thread at 8 warps per sub-partition; "fresh" = operands produced by the previous step (no reuse-cache hits).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/regbw tools/ubench/regbw.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
enum { ADD2_1FRESH, ADD2_2FRESH, FMA2_1FRESH, FMA2_2FRESH, FMA2_3FRESH, ADD1_2FRESH, FMA1_3FRESH, FMA1_1FRESH, NV };
static const char* names[NV] = {"FADD2  d = d + k            (1 fresh pair)", "FADD2  d = d + e            (2 fresh pairs)", "FFMA2  d = d * k + k2      (1 fresh pair)",
                                "FFMA2  d = e * k + d       (2 fresh pairs)", "FFMA2  d = e * f + d       (3 fresh pairs)", "FADD   d = d + e            (2 fresh)",
                                "FFMA   d = e * f + d       (3 fresh)", "FFMA   d = d * k + k2      (1 fresh)"};
template <int V>
__global__ void __launch_bounds__(1024) k(float* out, float seed, int iters, long long* cyc) {
  u64 p[8], q[8], r[8];
  float s[8], t[8], u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s[i] = seed + i + threadIdx.x; t[i] = seed * 0.5f + i; u[i] = 1.0f + i * 1e-3f;
    p[i] = ((u64)__float_as_uint(s[i]) << 32) | __float_as_uint(t[i]); q[i] = p[i] + 12345; r[i] = p[i] ^ 0x1111;
  }
  const u64 kk = ((u64)__float_as_uint(0.999f) << 32) | __float_as_uint(1.001f), k2 = ((u64)__float_as_uint(0.25f) << 32) | __float_as_uint(0.125f);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = (i + 1) & 7, l = (i + 3) & 7;
        if (V == ADD2_1FRESH) p[i] = add2(p[i], kk);
        if (V == ADD2_2FRESH) p[i] = add2(p[i], p[j]);
        if (V == FMA2_1FRESH) p[i] = fma2(p[i], kk, k2);
        if (V == FMA2_2FRESH) p[i] = fma2(p[j], kk, p[i]);
        if (V == FMA2_3FRESH) p[i] = fma2(p[j], p[l], p[i]);
        if (V == ADD1_2FRESH) s[i] = add1(s[i], s[j]);
        if (V == FMA1_3FRESH) s[i] = fma1(s[j], s[l], s[i]);
        if (V == FMA1_1FRESH) s[i] = fma1(s[i], 0.999f, 0.25f);
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += s[i] + t[i] + u[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)q[i]) + __uint_as_float((unsigned)r[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V>
void run() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
  const int iters = 2000, warps = 8;
  for (int rep = 0; rep < 2; ++rep) k<V><<<148, 128 * warps>>>(out, 1.0f, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-50s %5.2f cycles per warp-instruction on one sub-partition\n", names[V], (double)h / iters / 32.0 / warps);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<ADD2_1FRESH>(); run<ADD2_2FRESH>(); run<FMA2_1FRESH>(); run<FMA2_2FRESH>(); run<FMA2_3FRESH>(); run<ADD1_2FRESH>(); run<FMA1_3FRESH>(); run<FMA1_1FRESH>();
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
