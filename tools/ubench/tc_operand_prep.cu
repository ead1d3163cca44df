#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
// operand preparation of a tensor-core (tcgen05, FP16 x 3) 1024-point transform, WITHOUT the MMAs:
// what the CUDA cores still have to do per frame and stage.  warp = frame, lane = operand row.
__device__ __forceinline__ void split16(float re, float im, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(re, im);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(re - hf.x, im - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// stage 1: lane = n2 owns x[32 n1 + n2]; multiply by the window, split, write its own 128-byte row (hi) and (lo)
template <bool WIN>
__device__ __forceinline__ void prep_stage1(const float2* __restrict__ x, const float2* __restrict__ w, char* smem_hi, char* smem_lo, int lane) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float2 v = x[32 * i + lane];
    if (WIN) { const float2 ww = w[32 * i + lane]; v = make_float2(v.x * ww.x - v.y * ww.y, v.x * ww.y + v.y * ww.x); }
    split16(v.x, v.y, hi[i], lo[i]);
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int off = lane * 128 + ((c ^ (lane & 7)) << 4);
    *reinterpret_cast<uint4*>(smem_hi + off) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
    *reinterpret_cast<uint4*>(smem_lo + off) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
  }
}
// stage 2: lane = n2 holds Y[k1][n2] (32 values, as tcgen05.ld leaves them); twiddle, split, write element n2 of row k1
__device__ __forceinline__ void prep_stage2(const float2 (&y)[32], const float2* __restrict__ tw, char* smem_hi, char* smem_lo, int lane) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const float2 t = tw[k * 32 + lane];
    const float2 v = make_float2(y[k].x * t.x - y[k].y * t.y, y[k].x * t.y + y[k].y * t.x);
    uint32_t hi, lo;
    split16(v.x, v.y, hi, lo);
    const int off = k * 128 + (((lane >> 2) ^ (k & 7)) << 4) + ((lane & 3) << 2);
    *reinterpret_cast<uint32_t*>(smem_hi + off) = hi;
    *reinterpret_cast<uint32_t*>(smem_lo + off) = lo;
  }
}
extern __shared__ __align__(128) char smem[];
// MODE 1: stage-1 prep with window (transform B); 2: stage-1 without (transform A); 3: stage-2 prep (values from smem, as a stand-in for tcgen05.ld)
template <int MODE>
__global__ void __launch_bounds__(256, 1) prep_kernel(const float2* __restrict__ x, const float2* __restrict__ w, const float2* __restrict__ tw, int frames, int reps, float* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  char* hi = smem + warp * 8192;
  char* lo = hi + 4096;
  float acc = 0.f;
  for (int r = 0; r < reps; ++r)
    for (int f = blockIdx.x * 8 + warp; f < frames; f += gridDim.x * 8) {
      if (MODE == 1) prep_stage1<true>(x + (size_t)f * 1024, w, hi, lo, lane);
      if (MODE == 2) prep_stage1<false>(x + (size_t)f * 1024, w, hi, lo, lane);
      if (MODE == 3) {
        float2 y[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) y[k] = x[(size_t)f * 1024 + 32 * k + lane];
        prep_stage2(y, tw, hi, lo, lane);
      }
      __syncwarp();
      acc += *reinterpret_cast<float*>(hi + lane * 4);
    }
  if (acc == 123.456f) *sink = acc;
}

// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o build/tc_operand_prep tools/ubench/tc_operand_prep.cu
int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int frames = sms * 8 * 4, reps = 64;           // 38 MB of input: stays in L2, so the number is dispatch, not HBM
  float2 *x, *w, *tw;
  float* sink;
  cudaMalloc(&x, (size_t)frames * 1024 * sizeof(float2));
  cudaMalloc(&w, 1024 * sizeof(float2));
  cudaMalloc(&tw, 1024 * sizeof(float2));
  cudaMalloc(&sink, 4);
  cudaMemset(x, 0x3c, (size_t)frames * 1024 * sizeof(float2));
  cudaMemset(w, 0x3c, 1024 * sizeof(float2));
  cudaMemset(tw, 0x3c, 1024 * sizeof(float2));
  const size_t smem = 8 * 8192;
  cudaFuncSetAttribute(prep_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(prep_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(prep_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[3] = {"stage 1, window + split + 16 STS.128 (transform B)", "stage 1, split + 16 STS.128 (transform A)",
                          "stage 2, twiddle + split + 64 STS.32"};
  for (int mode = 1; mode <= 3; ++mode) {
    for (int it = 0; it < 2; ++it) {
      cudaEventRecord(e0);
      if (mode == 1) prep_kernel<1><<<sms, 256, smem>>>(x, w, tw, frames, reps, sink);
      if (mode == 2) prep_kernel<2><<<sms, 256, smem>>>(x, w, tw, frames, reps, sink);
      if (mode == 3) prep_kernel<3><<<sms, 256, smem>>>(x, w, tw, frames, reps, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      const cudaError_t le = cudaGetLastError();
      if (le != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(le)); return 1; }
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fr = (double)frames * reps;
    // 8 warps per SM = 2 per sub-partition: cycles one sub-partition spends per frame it processes
    const double cyc = ms * 1e-3 * (double)khz * 1e3 * sms * 4 / fr;
    printf("mode %d  %-52s %8.3f ms  %7.1f Gsamples/s-equivalent  %6.0f cycles per frame and sub-partition (nominal %d MHz)\n", mode, names[mode - 1], ms,
           fr * 1024 / (ms * 1e-3) / 1e9, cyc, khz / 1000);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
