// common.cuh — shared device helpers for the aether_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "aether_b200 kernels are written for sm_100a (B200) only"
#endif

namespace ae {

// device-side deferred error flags (reported by ae_sync)
enum : int { DEVERR_MOD_INDEX = 1 };

// ---- exact (never contracted) cf32 arithmetic: reproduces rustc/num-complex 0.2 ------------
// src/vecops.rs:94-155 via num-complex: every * + - / is a separately rounded IEEE op.
__device__ __forceinline__ float2 cx_scale_exact(float2 a, float s) {
  return make_float2(__fmul_rn(a.x, s), __fmul_rn(a.y, s));
}
__device__ __forceinline__ float2 cx_mul_exact(float2 a, float2 b) {
  return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                     __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cx_div_exact(float2 a, float2 b) {
  const float n = __fadd_rn(__fmul_rn(b.x, b.x), __fmul_rn(b.y, b.y));
  const float re = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
  const float im = __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
  return make_float2(__fdiv_rn(re, n), __fdiv_rn(im, n));
}
__device__ __forceinline__ float2 cx_add_exact(float2 a, float2 b) {
  return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y));
}
__device__ __forceinline__ float2 cx_sub_exact(float2 a, float2 b) {
  return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y));
}

// ---- fast cf32 arithmetic (FMA allowed) for the EVM-tier kernels (FFT, FIR) ----------------
__device__ __forceinline__ float2 cx_mul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cx_mul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
#ifdef AE_PACKED_F32X2
// sm_100a packed FP32: one FADD2 adds both halves of a cf32 (same FLOP rate, half the issue slots;
// measured in tools/ubench/f32x2.cu)
__device__ __forceinline__ unsigned long long cx_bits(float2 a) {
  return ((unsigned long long)__float_as_uint(a.y) << 32) | __float_as_uint(a.x);
}
__device__ __forceinline__ float2 cx_from_bits(unsigned long long v) {
  return make_float2(__uint_as_float((unsigned)v), __uint_as_float((unsigned)(v >> 32)));
}
__device__ __forceinline__ float2 cx_add(float2 a, float2 b) {
  unsigned long long r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b)));
  return cx_from_bits(r);
}
__device__ __forceinline__ float2 cx_sub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.f32x2 %0, %1, %2;" : "=l"(r) : "l"(cx_bits(a)), "l"(cx_bits(b)));
  return cx_from_bits(r);
}
#else
__device__ __forceinline__ float2 cx_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 cx_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ void cx_fma(float2& acc, float2 a, float2 b) {  // acc += a*b
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.x = fmaf(-a.y, b.y, acc.x);
  acc.y = fmaf(a.x, b.y, acc.y);
  acc.y = fmaf(a.y, b.x, acc.y);
}

// ---- streaming global access (data is touched once: do not pollute L1, evict-first in L2) --
__device__ __forceinline__ float2 ld_stream(const float2* p) { return __ldcs(p); }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float2* p, float2 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }

// QPSK / generic hard decision, exactly the float expressions of src/modulation.rs:36-49 and
// :133-140 with Iterator::min_by(partial_cmp.unwrap_or(Greater)) semantics: the later
// candidate wins iff best > later or the comparison is unordered.
template <int M>
__device__ __forceinline__ unsigned demod_index(float2 s, const float2 (&tab)[M]) {
  float d[M];
#pragma unroll
  for (int c = 0; c < M; ++c) {
    const float dr = __fsub_rn(s.x, tab[c].x), di = __fsub_rn(s.y, tab[c].y);
    d[c] = __fadd_rn(__fmul_rn(dr, dr), __fmul_rn(di, di));
  }
  unsigned best = 0;
  float bd = d[0];
#pragma unroll
  for (int c = 1; c < M; ++c)
    if (!(bd <= d[c])) { bd = d[c]; best = c; }
  return best;
}

// Hard decision for the GENERIC QPSK table (+-1, +-1) (src/modulation.rs:87-92), bit-identical to
// demod_index<4> but ~3x cheaper.  With A(+-) = fl(fl(re -+ 1)^2), B(+-) likewise, the reference
// compares d = fl(A + B).  Rounding is monotone, so the sign-based index is the reference's answer
// unless two candidates TIE after rounding.  A tie on the re decision needs
//     4|re| - 6.2u (|re|+1)^2  <=  2.01 u ((|re|+1)^2 + (|im|+1)^2),   u = 2^-24,
// which cannot happen when |re| > 2^-21 (1 + max(|re|,|im|))^2 (2x margin); same for im.  NaN, inf
// and the rare near-axis symbols fail the test (comparisons with NaN are false) and take the
// exact path, so ties ("first minimum wins") and NaN ("later index wins") are reproduced.
static __device__ __noinline__ unsigned demod_qpsk_exact_slow(float2 s) {  // one out-of-line copy: rare path
  const float2 tab[4] = {make_float2(1.0f, 1.0f), make_float2(-1.0f, 1.0f), make_float2(1.0f, -1.0f), make_float2(-1.0f, -1.0f)};
  return demod_index<4>(s, tab);
}
__device__ __forceinline__ unsigned demod_qpsk_generic(float2 s) {
  const float ax = fabsf(s.x), ay = fabsf(s.y);
  const float u = fmaf(fmaxf(ax, ay), 6.9053396600248786e-4f, 6.9053396600248786e-4f);  // 2^-10.5 (1 + max)
  const float thr = u * u;
  if (ax > thr && ay > thr) return (__float_as_uint(s.x) >> 31) | ((__float_as_uint(s.y) >> 31) << 1);
  return demod_qpsk_exact_slow(s);
}

// Same decision, split for hot loops: qpsk_fast_ok() is the tie-free test, qpsk_sign_index() the
// sign-based index; callers batch the rare failures and re-decide them with demod_qpsk_exact_slow().
__device__ __forceinline__ bool qpsk_fast_ok(float2 s) {
  const float ax = fabsf(s.x), ay = fabsf(s.y);
  const float u = fmaf(fmaxf(ax, ay), 6.9053396600248786e-4f, 6.9053396600248786e-4f);
  const float thr = u * u;
  return ax > thr && ay > thr;
}
// the two output bytes of one QPSK symbol as a little-endian u16: byte0 = idx & 1,
// byte1 = idx & 2 (compat=reference, hi_shift = 9) or (idx >> 1) & 1 (corrected, hi_shift = 8)
__device__ __forceinline__ unsigned qpsk_pair_from_signs(float2 s, unsigned hi_shift) {
  return (__float_as_uint(s.x) >> 31) | ((__float_as_uint(s.y) >> 31) << hi_shift);
}
__device__ __forceinline__ unsigned qpsk_pair_from_index(unsigned idx, unsigned hi_shift) {
  return (idx & 1u) | ((idx >> 1) << hi_shift);
}

// Philox4x32-10 (Salmon et al., Random123).  Known-answer vectors in tests/test_noise.py.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The ten round keys of a Philox4x32-10 stream depend only on the seed: the host expands them once
// (make_philox_keys) and kernels take them as a __grid_constant__ parameter, so the XORs read them
// straight from the constant bank instead of every thread re-deriving 20 key words per block.
struct PhiloxKeys { uint32_t k0[10], k1[10]; };
__host__ __device__ inline PhiloxKeys make_philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) { k.k0[r] = a; k.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
  return k;
}
__device__ __forceinline__ void philox4x32_10_keys(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& k, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k.k0[r];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k.k1[r];
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float sfu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Box-Muller on two 32-bit words -> two independent N(0,1) f32.
// AWGN is validated statistically (tier T2), so the transcendental part uses the SFU:
//   log : MUFU.LG2 (u1 >= 2^-33 is never subnormal), except within 2^-6 of 1 where its absolute error
//         would dominate the tiny result: there log(1 - t) = -t(1 + t/2 + t^2/3 + t^3/4) (t = 1 - u1 is exact);
//   sqrt: MUFU-based sqrt.approx (<= 2 ulp; exact 0 for u1 == 1);
//   trig: MUFU-based __sincosf on theta - pi in (-pi, pi], negated (cos/sin(theta) = -cos/sin(theta - pi)).
// |dz| stays below ~3e-6 sigma of the full-precision evaluation (tests compare with the oracle's).
__device__ __forceinline__ float2 gauss_pair(uint32_t a, uint32_t b) {
  const float u1 = __fmaf_rn(__uint2float_rn(a), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float u2 = __fmaf_rn(__uint2float_rn(b), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float t = 1.0f - u1;
  const float near1 = t * fmaf(t, fmaf(t, fmaf(t, 0.5f, 0.66666669f), 1.0f), 2.0f);   // -2 log(1 - t)
  const float m2l = t < 0.015625f ? near1 : -1.3862943611198906f * sfu_lg2(u1);          // -2 ln u1 = -2 ln2 lg2 u1
  const float r = sfu_sqrt(m2l);
  float s, c;
  __sincosf(fmaf(u2, 6.28318530717958647692f, -3.14159265358979323846f), &s, &c);
  return make_float2(-r * c, -r * s);
}

// noise pair for Philox block `pair` of stream (keys, stream): z0 = samples 2*pair, z1 = 2*pair+1
__device__ __forceinline__ void awgn_unit_pair(const PhiloxKeys& keys, uint64_t stream, uint64_t pair, float2& z0, float2& z1) {
  uint32_t o[4];
  philox4x32_10_keys((uint32_t)pair, (uint32_t)(pair >> 32), (uint32_t)stream, (uint32_t)(stream >> 32), keys, o);
  z0 = gauss_pair(o[0], o[1]);
  z1 = gauss_pair(o[2], o[3]);
}

// 256-bit streaming store (sm_100: STG.256): 32 bytes per lane, so a warp store covers 1 KiB contiguously.
// p must be 32-byte aligned.
__device__ __forceinline__ void st_stream_256(void* p, float4 a, float4 b) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
               "f"(b.z), "f"(b.w)
               : "memory");
}

// 256-bit streaming load (LDG.256); p must be 32-byte aligned
__device__ __forceinline__ void ld_stream_256(const void* p, float4& a, float4& b) {
  asm volatile("ld.global.cs.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

}  // namespace ae
