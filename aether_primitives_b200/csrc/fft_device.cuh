// fft_device.cuh — in-register / shared-memory Stockham FFT for one power-of-two frame.
//
// Replaces the rustfft call under Cfft::{fwd,bwd,ifwd,ibwd,tfwd,tbwd} (src/fft.rs:162-230).
//
// Layout.  A frame of N = 2^L points is owned by T = N/16 threads, 16 points per thread.
// Register m of thread t always holds the element at position  t + m*T  — on input of EVERY
// pass and on output of the LAST pass.  So global loads/stores are coalesced 8-byte accesses
// (consecutive t -> consecutive cf32), the result is in natural order, and two transforms can
// be chained in registers (FFT -> pointwise multiply -> inverse FFT) with no exchange between.
//
// Passes.  N = 16^A * 2^REM  ->  A radix-16 passes then one radix-2^REM pass (REM = L%4).
// Pass p with sub-transform size NS = 16^p and radix R:
//     butterfly j = t + q*T (q < 16/R),  k = j mod NS
//     v[r]  = x[q + (16/R) r] * W_N^( r*k*N/(NS*R) )         (twiddle; none when NS == 1)
//     v     = DFT_R(v)
//     out position = (j - k)*R + k + r*NS                     (Stockham autosort)
// Between passes the 16 values go through shared memory (one cf32 of padding per 16 so the
// stride-16 stores of the first pass are bank-conflict free); the last pass leaves them in
// registers.  Twiddles come from a table tw[k] = exp(-2 pi i k/N) computed in f64 on the host
// and rounded to f32 (same accuracy class as rustfft), read through the read-only path.
#pragma once
#include "common.cuh"

namespace ae {

constexpr int ilog2_c(unsigned long long v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

template <int N>
struct FftCfg {
  static constexpr int L = ilog2_c(N);
  static constexpr int A = L / 4;          // radix-16 passes
  static constexpr int REM = L % 4;        // last pass radix 2^REM
  static constexpr int NP = A + (REM ? 1 : 0);
  static constexpr int T = N / 16;         // threads per frame
  static constexpr int SMEM_ELEMS = N + N / 16;  // padded cf32 per frame
  static_assert(N >= 16 && (N & (N - 1)) == 0, "power of two >= 16");
  static constexpr int radix(int p) { return p < A ? 16 : (1 << REM); }
  static constexpr int ns(int p) { return 1 << (4 * p); }  // 16^p
};

__device__ __forceinline__ int fft_pad(int i) { return i + (i >> 4); }

// multiply by -i (forward) / +i (inverse)
template <bool INV>
__device__ __forceinline__ float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a * (wr, -wi) forward, a * (wr, +wi) inverse; (wr, wi) = (cos, sin) of the angle magnitude
template <bool INV>
__device__ __forceinline__ float2 mul_w(float2 a, float wr, float wi) {
  return INV ? make_float2(fmaf(a.x, wr, -a.y * wi), fmaf(a.x, wi, a.y * wr))
             : make_float2(fmaf(a.x, wr, a.y * wi), fmaf(a.y, wr, -a.x * wi));
}
// table twiddle: tw holds exp(-i theta); inverse uses the conjugate
template <bool INV>
__device__ __forceinline__ float2 mul_tw(float2 a, float2 w) {
  return INV ? cx_mul_conj(a, w) : cx_mul(a, w);
}

template <bool INV>
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 s = cx_add(a, b), d = cx_sub(a, b);
  a = s; b = d;
}
template <bool INV>
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s0 = cx_add(a0, a2), s1 = cx_sub(a0, a2);
  const float2 s2 = cx_add(a1, a3), s3 = mul_mi<INV>(cx_sub(a1, a3));
  a0 = cx_add(s0, s2); a1 = cx_add(s1, s3); a2 = cx_sub(s0, s2); a3 = cx_sub(s1, s3);
}

constexpr float kC8 = 0.70710678118654752440f;   // cos(pi/4)
constexpr float kC16 = 0.92387953251128675613f;  // cos(pi/8)
constexpr float kS16 = 0.38268343236508977173f;  // sin(pi/8)

// natural-order in, natural-order out DFTs on register arrays
template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<1, INV> {
  static __device__ __forceinline__ void run(float2 (&)[1]) {}
};
template <bool INV> struct Dft<2, INV> {
  static __device__ __forceinline__ void run(float2 (&v)[2]) { dft2<INV>(v[0], v[1]); }
};
template <bool INV> struct Dft<4, INV> {
  static __device__ __forceinline__ void run(float2 (&v)[4]) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV> struct Dft<8, INV> {
  // n = 2a + b, k = c + 4d:  W8^(nk) = W4^(ac) * W8^(bc) * W2^(bd)
  static __device__ __forceinline__ void run(float2 (&v)[8]) {
    dft4<INV>(v[0], v[2], v[4], v[6]);  // b = 0 -> Y0[c] in v[2c]
    dft4<INV>(v[1], v[3], v[5], v[7]);  // b = 1 -> Y1[c] in v[2c+1]
    v[3] = mul_w<INV>(v[3], kC8, kC8);                 // W8^1
    v[5] = mul_mi<INV>(v[5]);                          // W8^2 = -i
    v[7] = mul_w<INV>(v[7], -kC8, kC8);                // W8^3
    float2 o[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      o[c] = cx_add(v[2 * c], v[2 * c + 1]);
      o[c + 4] = cx_sub(v[2 * c], v[2 * c + 1]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = o[i];
  }
};
template <bool INV> struct Dft<16, INV> {
  // n = 4a + b, k = c + 4d:  W16^(nk) = W4^(ac) * W16^(bc) * W4^(bd)
  static __device__ __forceinline__ void run(float2 (&v)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);  // Y_b[c] in v[4c+b]
    // W16^(b*c), b,c in 1..3
    v[5] = mul_w<INV>(v[5], kC16, kS16);     // e=1
    v[6] = mul_w<INV>(v[6], kC8, kC8);       // e=2
    v[7] = mul_w<INV>(v[7], kS16, kC16);     // e=3
    v[9] = mul_w<INV>(v[9], kC8, kC8);       // e=2
    v[10] = mul_mi<INV>(v[10]);              // e=4
    v[11] = mul_w<INV>(v[11], -kC8, kC8);    // e=6
    v[13] = mul_w<INV>(v[13], kS16, kC16);   // e=3
    v[14] = mul_w<INV>(v[14], -kC8, kC8);    // e=6
    v[15] = mul_w<INV>(v[15], -kC16, -kS16); // e=9: cos=-c16, sin(2pi*9/16) = -s16
#pragma unroll
    for (int c = 0; c < 4; ++c) dft4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);  // X[c+4d] in v[4c+d]
    float2 o[16];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int d = 0; d < 4; ++d) o[c + 4 * d] = v[4 * c + d];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = o[i];
  }
};

// one pass; LAST leaves the result in x (register m <-> position t + m*T), otherwise it is
// written to the padded shared-memory frame `sm`
template <int N, int P, bool INV>
__device__ __forceinline__ void fft_pass(float2 (&x)[16], float2* __restrict__ sm, const float2* __restrict__ tw, int t) {
  using C = FftCfg<N>;
  constexpr int R = C::radix(P);
  constexpr int NS = C::ns(P);
  constexpr int B = 16 / R;
  constexpr int T = C::T;
  constexpr bool LAST = (P == C::NP - 1);
  constexpr int TWS = N / (NS * R);
#pragma unroll
  for (int q = 0; q < B; ++q) {
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = x[q + B * r];
    const int j = t + q * T;
    const int k = j & (NS - 1);
    if (NS > 1) {
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = mul_tw<INV>(v[r], __ldg(tw + r * k * TWS));
    }
    Dft<R, INV>::run(v);
    if (LAST) {
#pragma unroll
      for (int r = 0; r < R; ++r) x[q + B * r] = v[r];
    } else {
      const int base = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) sm[fft_pad(base + r * NS)] = v[r];
    }
  }
}

template <int N>
__device__ __forceinline__ void fft_load_smem(float2 (&x)[16], const float2* __restrict__ sm, int t) {
  constexpr int T = FftCfg<N>::T;
#pragma unroll
  for (int m = 0; m < 16; ++m) x[m] = sm[fft_pad(t + m * T)];
}

// barrier among the T threads that own one frame (frame slot f of the CTA)
template <int T>
__device__ __forceinline__ void frame_sync(int f) {
  if (T < 32) __syncwarp();
  else if (T == (int)blockDim.x) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(f + 1), "r"(T) : "memory");
}

template <int N, int P, bool INV>
__device__ __forceinline__ void fft_passes_from(float2 (&x)[16], float2* __restrict__ sm, const float2* __restrict__ tw, int t, int f) {
  using C = FftCfg<N>;
  fft_pass<N, P, INV>(x, sm, tw, t);
  if constexpr (P + 1 < C::NP) {
    frame_sync<C::T>(f);
    fft_load_smem<N>(x, sm, t);
    if constexpr (P + 2 < C::NP) frame_sync<C::T>(f);  // WAR: the next pass stores into sm again
    fft_passes_from<N, P + 1, INV>(x, sm, tw, t, f);
  }
}

// Full transform of one frame.  x: register m <-> position t + m*T (in and out).
// The caller must make sure every thread of the frame is past its last shared-memory READ of a
// previous use of `sm` (frame_sync) before calling this again with the same buffer.
template <int N, bool INV>
__device__ __forceinline__ void fft_frame(float2 (&x)[16], float2* __restrict__ sm, const float2* __restrict__ tw, int t, int f) {
  fft_passes_from<N, 0, INV>(x, sm, tw, t, f);
}

}  // namespace ae
