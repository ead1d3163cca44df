// Micro-benchmark: throughput of the warp-level (legacy) mma.sync path on sm_100a, bf16 m16n8k16 and
// tf32 m16n8k8 with fp32 accumulation.  Answers whether a DFT-as-GEMM stage can live on mma.sync or
// needs tcgen05.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE, int NACC>
__global__ void __launch_bounds__(256) k(float* out, unsigned seed, int iters) {
  float d[NACC][4];
  unsigned a[4], b[2];
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) a[j] = seed * (threadIdx.x + j + 1);
  b[0] = seed ^ threadIdx.x;
  b[1] = seed + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0) mma_bf16(d[i], a, b);
      else mma_tf32(d[i], a, b);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE, int NACC>
void run(const char* name, int ctas_per_sm) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
  const int iters = 8192;
  k<MODE, NACC><<<148 * ctas_per_sm, 256>>>(out, 0, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, NACC><<<148 * ctas_per_sm, 256>>>(out, 0, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop_per_mma = MODE == 0 ? 2.0 * 16 * 8 * 16 : 2.0 * 16 * 8 * 8;
  const double flops = flop_per_mma * NACC * (double)iters * 8 /*warps*/ * 148 * ctas_per_sm;
  printf("%-28s NACC=%d ctas/SM=%d  %.3f ms  %.1f TFLOP/s\n", name, NACC, ctas_per_sm, ms, flops / ms * 1e-9);
  cudaFree(out);
}

int main() {
  run<0, 4>("mma.sync bf16 m16n8k16", 1);
  run<0, 8>("mma.sync bf16 m16n8k16", 1);
  run<0, 8>("mma.sync bf16 m16n8k16", 2);
  run<0, 16>("mma.sync bf16 m16n8k16", 2);
  run<1, 4>("mma.sync tf32 m16n8k8", 1);
  run<1, 8>("mma.sync tf32 m16n8k8", 1);
  run<1, 8>("mma.sync tf32 m16n8k8", 2);
  run<1, 16>("mma.sync tf32 m16n8k8", 2);
  return cudaDeviceSynchronize() != cudaSuccess;
}
