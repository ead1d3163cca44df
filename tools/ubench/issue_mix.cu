// Micro-benchmark: does a packed FFMA2 occupy the SMSP's issue port for one cycle or for two?
// Each mode runs independent instruction chains per thread; reported: SM cycles per loop iteration per SMSP
// at W warps per SMSP (throughput regime).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned d; asm volatile("lop3.b32 %0, %1, %2, %1, 0x1e;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ u64 pack(float x, float y) { return ((u64)__float_as_uint(y) << 32) | __float_as_uint(x); }

// NP packed FMAs, NS scalar FMAs, NA ALU ops per iteration, all independent chains
template <int NP, int NS, int NA>
__global__ void __launch_bounds__(1024) k(float* out, float a0, int iters, long long* cyc) {
  u64 p[8]; float s[16]; unsigned q[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pack(a0 + i + threadIdx.x, a0 - i);
#pragma unroll
  for (int i = 0; i < 16; ++i) { s[i] = a0 + i + threadIdx.x; q[i] = threadIdx.x * 7 + i; }
  const float bf = 0.999f + threadIdx.x * 1e-9f;
  const u64 b = pack(bf, bf), c = pack(bf * 0.25f, bf * 0.125f);
  const unsigned m = threadIdx.x | 0x5a5a0000u;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < NP) p[i] = fma2(p[i], b, c);
        if (2 * i < NS) s[2 * i] = fma1(s[2 * i], bf, 0.3f);
        if (2 * i + 1 < NS) s[2 * i + 1] = fma1(s[2 * i + 1], bf, 0.3f);
        if (2 * i < NA) q[2 * i] = lop(q[2 * i], q[(2 * i + 5) & 15] + m);
        if (2 * i + 1 < NA) q[2 * i + 1] = lop(q[2 * i + 1], q[(2 * i + 6) & 15] + m);
      }
    }
  }
  const long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
#pragma unroll
  for (int i = 0; i < 16; ++i) r += s[i] + (float)q[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NP, int NS, int NA>
void run(const char* name) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps_per_smsp : {1, 4, 8}) {
    const int threads = 128 * warps_per_smsp;
    k<NP, NS, NA><<<148, threads>>>(out, 1.0f, iters, cyc);
    k<NP, NS, NA><<<148, threads>>>(out, 1.0f, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_iter = (double)h / iters / 4.0;   // cycles per (NP packed + NS scalar + NA alu) group, all warps of an SMSP together
    printf("%-34s warps/SMSP=%d  %.2f cycles per group per SMSP  (%.2f per warp-group)  instr/group=%d\n", name, warps_per_smsp, per_iter,
           per_iter / warps_per_smsp, NP + NS + NA);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<8, 0, 0>("8 FFMA2");
  run<0, 16, 0>("16 FFMA");
  run<0, 0, 16>("16 LOP");
  run<8, 0, 8>("8 FFMA2 + 8 LOP");
  run<8, 0, 16>("8 FFMA2 + 16 LOP");
  run<8, 8, 0>("8 FFMA2 + 8 FFMA");
  run<8, 16, 0>("8 FFMA2 + 16 FFMA");
  run<0, 16, 8>("16 FFMA + 8 LOP");
  run<0, 16, 16>("16 FFMA + 16 LOP");
  run<4, 8, 8>("4 FFMA2 + 8 FFMA + 8 LOP");
  return 0;
}
