// Micro-benchmark: dispatch cost of single instructions on one sub-partition (cycles per warp-instruction at 8 warps per
// sub-partition, 8 independent chains per thread), for the cost model of DESIGN.md 5.2.2 / 9.2: what does a Philox round
// cost, what do the Box-Muller pieces cost.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/issue_cost tools/ubench/issue_cost.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned u32;

enum Op { IMAD_WIDE, IMAD_LO, IMAD_HI, LOP3_XOR3, IADD3, SHF, PRMT, POPC, I2FP, FFMA, FMUL, MUFU_LG2, MUFU_SIN, MUFU_SQRT, MUFU_EX2, FMNMX3, FSETP_FSEL,
          PHILOX_WIDE, PHILOX_HILO, HMMA_TF32, FADD2, LDS64, LDS128, STS32, NOPS };
static const char* kNames[NOPS] = {"IMAD.WIDE.U32 (mul.wide.u32)", "IMAD (mul.lo.u32)", "IMAD.HI (mul.hi.u32)", "LOP3 (a ^ b ^ c)", "IADD3", "SHF (funnel shift)",
                                   "PRMT", "POPC", "I2FP.F32.U32", "FFMA", "FMUL", "MUFU.LG2", "MUFU.SIN", "MUFU.SQRT (sqrt.approx)", "MUFU.EX2", "FMNMX3 (3-input min)",
                                   "FSETP + FSEL", "Philox round, mul.wide form (2 IMAD.WIDE + 2 LOP3)", "Philox round, mul.lo + mul.hi form (4 IMAD + 2 LOP3)", "HMMA.1688.F32.TF32 (mma.sync m16n8k8)", "FADD2 (add.f32x2)",
                                   "LDS.64 (conflict-free)", "LDS.128 (conflict-free)", "STS.32 (conflict-free)"};

template <int OP>
__device__ __forceinline__ void step(u32& a, u32& b, u32& c, u32& d, u32 k) {
  if (OP == IMAD_WIDE) { u64 p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a), "r"(0xD2511F53u)); a = (u32)p ^ 0; b = (u32)(p >> 32); }
  if (OP == IMAD_LO) asm volatile("mul.lo.u32 %0, %0, %1;" : "+r"(a) : "r"(0xD2511F53u));
  if (OP == IMAD_HI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a) : "r"(0xD2511F53u));
  if (OP == LOP3_XOR3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(k));
  if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(k));
  if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(b));
  if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(a) : "r"(b));
  if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a));
  if (OP == I2FP) { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a)); a = __float_as_uint(f); }
  if (OP == FFMA) { float f = __uint_as_float(a); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(0.999f), "f"(0.25f)); a = __float_as_uint(f); }
  if (OP == FMUL) { float f = __uint_as_float(a); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(0.999f)); a = __float_as_uint(f); }
  if (OP == MUFU_LG2) { float f = __uint_as_float(a); asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f)); a = __float_as_uint(f); }
  if (OP == MUFU_SIN) { float f = __uint_as_float(a); asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f)); a = __float_as_uint(f); }
  if (OP == MUFU_SQRT) { float f = __uint_as_float(a); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f)); a = __float_as_uint(f); }
  if (OP == MUFU_EX2) { float f = __uint_as_float(a); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f)); a = __float_as_uint(f); }
  if (OP == FMNMX3) { float f = __uint_as_float(a); asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); a = __float_as_uint(f); }
  if (OP == FSETP_FSEL) { float f = __uint_as_float(a); f = f < 0.015625f ? __uint_as_float(b) : __uint_as_float(c); a = __float_as_uint(f) + 1; }
  if (OP == PHILOX_WIDE) {
    u64 p0, p1;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(a), "r"(0xD2511F53u));
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(c), "r"(0xCD9E8D57u));
    const u32 n0 = (u32)(p1 >> 32) ^ b ^ k, n2 = (u32)(p0 >> 32) ^ d ^ (k + 1);
    b = (u32)p1; d = (u32)p0; a = n0; c = n2;
  }
  if (OP == HMMA_TF32) {
    float e0 = __uint_as_float(a), e1 = __uint_as_float(b), e2 = __uint_as_float(c), e3 = __uint_as_float(d);
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(e0), "+f"(e1), "+f"(e2), "+f"(e3) : "r"(k), "r"(k + 1), "r"(k + 2), "r"(k + 3), "r"(k + 4), "r"(k + 5));
    a = __float_as_uint(e0); b = __float_as_uint(e1); c = __float_as_uint(e2); d = __float_as_uint(e3);
  }
  if (OP == FADD2) {
    u64 v = ((u64)b << 32) | a, w = ((u64)k << 32) | k;
    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(w));
    a = (u32)v; b = (u32)(v >> 32);
  }
  if (OP == LDS64 || OP == LDS128 || OP == STS32) {
    extern __shared__ __align__(16) unsigned char sm[];
    const u32 base = (u32)__cvta_generic_to_shared(sm) + threadIdx.x * (OP == LDS128 ? 16 : (OP == LDS64 ? 8 : 4));
    if (OP == LDS64) { u32 x, y; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(base + (a & 0x3000))); a ^= x; b += y; }
    if (OP == LDS128) { u32 x, y, z, w; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base + (a & 0x3000))); a ^= x ^ z; b += y + w; }
    if (OP == STS32) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + (k & 0x3000)), "r"(a) : "memory"); }
  }
  if (OP == PHILOX_HILO) {
    u32 l0, h0, l1, h1;
    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l0) : "r"(a), "r"(0xD2511F53u));
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h0) : "r"(a), "r"(0xD2511F53u));
    asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l1) : "r"(c), "r"(0xCD9E8D57u));
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h1) : "r"(c), "r"(0xCD9E8D57u));
    const u32 n0 = h1 ^ b ^ k, n2 = h0 ^ d ^ (k + 1);
    b = l1; d = l0; a = n0; c = n2;
  }
}

template <int OP>
__global__ void __launch_bounds__(1024) k(u32* out, u32 seed, int iters, long long* cyc) {
  u32 a[8], b[8], c[8], d[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 13 + i; b[i] = a[i] * 7 + 1; c[i] = a[i] ^ 0x3f800000u; d[i] = i + 0x40000000u; }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) step<OP>(a[i], b[i], c[i], d[i], seed + u);
  }
  const long long t1 = clock64();
  u32 r = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) r += a[i] + b[i] + c[i] + d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run() {
  u32* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * sizeof(u32)); cudaMalloc(&cyc, 8);
  const int iters = 1000, warps = 8;
  cudaFuncSetAttribute(k<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<OP><<<148, 128 * warps, 64 * 1024>>>(out, 12345u, iters, cyc);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // 32 steps per iteration per warp, `warps` warps share the sub-partition
  printf("%-62s %6.2f cycles per warp-step on one sub-partition\n", kNames[OP], (double)h / iters / 32.0 / warps);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<IMAD_WIDE>(); run<IMAD_LO>(); run<IMAD_HI>(); run<LOP3_XOR3>(); run<IADD3>(); run<SHF>(); run<PRMT>(); run<POPC>(); run<I2FP>(); run<FFMA>(); run<FMUL>();
  run<MUFU_LG2>(); run<MUFU_SIN>(); run<MUFU_SQRT>(); run<MUFU_EX2>(); run<FMNMX3>(); run<FSETP_FSEL>(); run<PHILOX_WIDE>(); run<PHILOX_HILO>();
  run<HMMA_TF32>(); run<FADD2>(); run<LDS64>(); run<LDS128>(); run<STS32>();
  printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
