// Micro-benchmark: issue rate of FP32 instruction forms on sm_100a (informs the FP32 ceilings in DESIGN.md).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma_forms tools/ubench/ffma_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

__constant__ float cbank[64];

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a0, float b0, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = a0 + i + threadIdx.x;
  float b = b0 + threadIdx.x * 1e-9f, c = b0 * 1.5f + threadIdx.x * 1e-9f;  // per-thread: real vector registers
  float xv[16], yv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { xv[i] = b + i * 1e-8f; yv[i] = c - i * 1e-8f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (MODE == 0) acc[i] = fmaf(acc[i], b, c);                 // FFMA R, R, R, R  (3 register sources)
        if (MODE == 1) acc[i] = fmaf(acc[i], cbank[(u * 16 + i) & 63], c);  // FFMA R, R, c[][], R
        if (MODE == 2) acc[i] = fmaf(acc[i], 1.0000001f, c);        // FFMA R, R, imm, R
        if (MODE == 3) acc[i] = acc[i] + b;                         // FADD R, R, R
        if (MODE == 4) acc[i] = acc[i] * b;                         // FMUL R, R, R
        if (MODE == 5) acc[i] = fmaf(b, c, acc[i]);                 // FFMA accumulate form: R_acc += b*c
        if (MODE == 6) acc[i] = fmaf(b, cbank[(u * 16 + i) & 63], acc[i]);  // accumulate with constant-bank operand
        if (MODE == 7) acc[i] = fmaf(xv[i], yv[(i + u) & 15], acc[i]);      // three distinct vector registers
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
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
  const double inst = (double)148 * 8 * 256 * iters * 8 * 16;  // thread-level instructions
  printf("%-34s %.3f ms  %.2f T thread-instr/s  = %.2f per clk per SM (of 128 lanes) at 1965 MHz\n", name, ms, inst / ms / 1e9,
         inst / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  float h[64];
  for (int i = 0; i < 64; ++i) h[i] = 1.0f + i * 1e-7f;
  cudaMemcpyToSymbol(cbank, h, sizeof(h));
  run<0>("FFMA acc = acc*b + c (R,R,R)");
  run<5>("FFMA acc = b*c + acc (R,R,R)");
  run<1>("FFMA acc = acc*const + c");
  run<6>("FFMA acc = b*const + acc");
  run<2>("FFMA acc = acc*imm + c");
  run<7>("FFMA acc += x[i]*y[j] (3 vregs)");
  run<3>("FADD acc = acc + b");
  run<4>("FMUL acc = acc * b");
  return 0;
}
