// fft_device.cuh — in-register / shared-memory Stockham FFT for power-of-two frames.
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
// registers.
//
// Twiddles exp(-2 pi i k/N) are computed in f64 on the host and rounded to f32 (rustfft's accuracy
// class).  Reading a plain table W[k] costs scattered 8-byte loads, and the first profile of the
// fused chain showed the L1 data pipe at 81 % because of them, so the host re-lays the values out
// PER THREAD (`tw` below is that table: row-major [row][t], coalesced, L1-resident) and
//   * a radix-16 pass loads only W^1,W^2,W^3 and W^4,W^8,W^12 and forms W^(4a+b) = W^(4a) W^b;
//   * the last (radix-2^REM) pass has k = t + q*T, so W^(r k) = W^(r t) * W16^(r q): R-1 loads,
//     the rest are compile-time constants;
//   * NB frames that share the thread mapping (the two transforms of the fused chain) are run
//     pass by pass TOGETHER so every twiddle is loaded once and every barrier is shared.
#pragma once
#include <type_traits>
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
  // per-thread twiddle table (built on the host, see fft_thread_twiddles): pass p >= 1 owns
  // tw_entries(p) rows of T values each, row-major [row][t] so a warp reads contiguous cf32
  static constexpr int tw_entries(int p) { return p == 0 ? 0 : (radix(p) == 16 ? 6 : radix(p) - 1); }
  static constexpr int tw_row0(int p) { return p <= 1 ? 0 : tw_row0(p - 1) + tw_entries(p - 1); }
  static constexpr int TW_ROWS = tw_row0(NP - 1) + tw_entries(NP - 1);
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

constexpr float kC8 = 0.70710678118654752440f;   // cos(pi/4)
constexpr float kC16 = 0.92387953251128675613f;  // cos(pi/8)
constexpr float kS16 = 0.38268343236508977173f;  // sin(pi/8)

// a * W16^E (forward) or its conjugate (inverse), E known at compile time
template <int E, bool INV>
__device__ __forceinline__ float2 mul_w16(float2 a) {
  constexpr int e = ((E % 16) + 16) % 16;
  if constexpr (e == 0) return a;
  else if constexpr (e == 4) return mul_mi<INV>(a);
  else if constexpr (e == 8) return make_float2(-a.x, -a.y);
  else if constexpr (e == 12) return mul_mi<!INV>(a);
  else {
    // cos / sin of 2 pi e / 16
    constexpr float c = (e == 1 || e == 15) ? kC16 : (e == 2 || e == 14) ? kC8 : (e == 3 || e == 13) ? kS16
                        : (e == 5 || e == 11) ? -kS16 : (e == 6 || e == 10) ? -kC8 : -kC16;  // e == 7 || e == 9
    constexpr float s = (e == 1 || e == 7) ? kS16 : (e == 2 || e == 6) ? kC8 : (e == 3 || e == 5) ? kC16
                        : (e == 9 || e == 15) ? -kS16 : (e == 10 || e == 14) ? -kC8 : -kC16;  // e == 11 || e == 13
    return mul_w<INV>(a, c, s);
  }
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
    v[3] = mul_w16<2, INV>(v[3]);       // W8^1
    v[5] = mul_w16<4, INV>(v[5]);       // W8^2 = -i
    v[7] = mul_w16<6, INV>(v[7]);       // W8^3
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
    v[5] = mul_w16<1, INV>(v[5]);
    v[6] = mul_w16<2, INV>(v[6]);
    v[7] = mul_w16<3, INV>(v[7]);
    v[9] = mul_w16<2, INV>(v[9]);
    v[10] = mul_w16<4, INV>(v[10]);
    v[11] = mul_w16<6, INV>(v[11]);
    v[13] = mul_w16<3, INV>(v[13]);
    v[14] = mul_w16<6, INV>(v[14]);
    v[15] = mul_w16<9, INV>(v[15]);
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

// compile-time loop helper: f(integral_constant<int, I>) for I in [0, N)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// one pass over NB frames that share thread mapping and twiddles.  LAST leaves the result in x
// (register m <-> position t + m*T), otherwise it is written to the padded shared frames sm[b].
//
// TAIL0: the caller only needs the LAST T bins (register 15) of frame 0.  Everything is fully
// unrolled register code, so the compiler removes whatever does not feed x[0][15] on its own (the
// other butterflies of the last pass, their twiddle products, the shared-memory loads); the one thing
// it cannot see is which STORES of the pass before the last are never read back: with
// N = 16*NS*Rlast, thread t' = a*NS + k writes output r to position NS*(16a + r) + k, i.e. to
// register m' = a*(16/Rlast) + floor(r/Rlast) of the last pass, and only m' = q + (16/Rlast) r' with
// q = 16/Rlast - 1 is consumed  =>  only r >= 16 - Rlast has to be stored.
template <int N, int P, bool INV, int NB, bool TAIL0 = false>
__device__ __forceinline__ void fft_pass(float2 (&x)[NB][16], float2* const (&sm)[NB], const float2* __restrict__ tw, int t) {
  using C = FftCfg<N>;
  constexpr int R = C::radix(P);
  constexpr int NS = C::ns(P);
  constexpr int B = 16 / R;
  constexpr int T = C::T;
  constexpr bool LAST = (P == C::NP - 1);
  constexpr int TWS = N / (NS * R);
  if constexpr (R == 16) {
    const int k = t & (NS - 1);
    (void)TWS;
    float2 v[NB][16];
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int r = 0; r < 16; ++r) v[b][r] = x[b][r];
    if constexpr (NS > 1) {
      // rows: W^u, W^2u, W^3u, W^4u, W^8u, W^12u with u = (t mod NS) * TWS, one value per thread
      const float2* row = tw + C::tw_row0(P) * T + t;
      float2 wl[4], wh[4];  // W^b (b = 1..3) and W^(4a) (a = 1..3)
#pragma unroll
      for (int i = 1; i < 4; ++i) { wl[i] = __ldg(row + (i - 1) * T); wh[i] = __ldg(row + (i + 2) * T); }
#pragma unroll
      for (int r = 1; r < 16; ++r) {
        const int a = r >> 2, bb = r & 3;
        float2 w;
        if (a == 0) w = wl[bb];
        else if (bb == 0) w = wh[a];
        else w = cx_mul(wh[a], wl[bb]);
#pragma unroll
        for (int b = 0; b < NB; ++b) v[b][r] = mul_tw<INV>(v[b][r], w);
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) Dft<16, INV>::run(v[b]);
    if constexpr (LAST) {
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < 16; ++r) x[b][r] = v[b][r];
    } else {
      // padded address of position (t-k)*16 + k + r*NS is linear in r
      int a0;
      constexpr int STEP = (NS == 1) ? 1 : (NS + NS / 16);
      if constexpr (NS == 1) a0 = 17 * t;
      else a0 = fft_pad((t - k) * 16 + k);
      constexpr bool BEFORE_LAST = (P == C::NP - 2);
      constexpr int RLAST = C::radix(C::NP - 1);
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          if (TAIL0 && b == 0 && BEFORE_LAST && r < 16 - RLAST) continue;  // never read back
          sm[b][a0 + r * STEP] = v[b][r];
        }
    }
  } else {
    // last pass, radix R < 16, B = 16/R butterflies per thread; k = j = t + q*T, TWS == 1
    static_assert(LAST && TWS == 1, "sub-radix pass must be the last one");
    float2 wt[R];  // W^(r t), rows r = 1..R-1 of this pass
    const float2* row = tw + C::tw_row0(P) * T + t;
#pragma unroll
    for (int r = 1; r < R; ++r) wt[r] = __ldg(row + (r - 1) * T);
    static_for<0, B>([&](auto qc) {
      constexpr int q = decltype(qc)::value;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float2 v[R];
        v[0] = x[b][q];
        static_for<1, R>([&](auto rc) {
          constexpr int r = decltype(rc)::value;
          // W^(r (t + q T)) = W^(r t) * W16^(r q)   (T = N/16)
          v[r] = mul_w16<r * q, INV>(mul_tw<INV>(x[b][q + B * r], wt[r]));
        });
        Dft<R, INV>::run(v);
#pragma unroll
        for (int r = 0; r < R; ++r) x[b][q + B * r] = v[r];
      }
    });
  }
}

template <int N, int NB>
__device__ __forceinline__ void fft_load_smem(float2 (&x)[NB][16], float2* const (&sm)[NB], int t) {
  constexpr int T = FftCfg<N>::T;
  if constexpr (T % 16 == 0) {
    const int a0 = fft_pad(t);
    constexpr int STEP = T + T / 16;
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int m = 0; m < 16; ++m) x[b][m] = sm[b][a0 + m * STEP];
  } else {
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int m = 0; m < 16; ++m) x[b][m] = sm[b][fft_pad(t + m * T)];
  }
}

// barrier among the T threads that own one frame (frame slot f of the CTA)
template <int T>
__device__ __forceinline__ void frame_sync(int f) {
  if (f < 0) { __syncthreads(); return; }  // column kernels: a frame's threads are spread over the CTA
  if (T < 32) __syncwarp();
  else if (T == (int)blockDim.x) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(f + 1), "r"(T) : "memory");
}

struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};
// `after_first_sync` runs once, right after the first slot barrier: at that point every thread of
// the slot has consumed its input registers (used to issue the next frame's TMA copy).
template <int N, int P, bool INV, int NB, bool TAIL0 = false, class Hook = NoHook>
__device__ __forceinline__ void fft_passes_from(float2 (&x)[NB][16], float2* const (&sm)[NB], const float2* __restrict__ tw, int t, int f,
                                                Hook after_first_sync = Hook()) {
  using C = FftCfg<N>;
  fft_pass<N, P, INV, NB, TAIL0>(x, sm, tw, t);
  if constexpr (P + 1 < C::NP) {
    frame_sync<C::T>(f);
    if constexpr (P == 0) after_first_sync();
    fft_load_smem<N, NB>(x, sm, t);
    if constexpr (P + 2 < C::NP) frame_sync<C::T>(f);  // WAR: the next pass stores into sm again
    fft_passes_from<N, P + 1, INV, NB, TAIL0, NoHook>(x, sm, tw, t, f);
  }
}

// Full transform of NB frames at once.  x[b]: register m <-> position t + m*T (in and out).
// The caller must make sure every thread of the frame is past its last shared-memory READ of a
// previous use of the buffers (frame_sync) before calling this again with the same buffers.
template <int N, bool INV, int NB, bool TAIL0 = false, class Hook = NoHook>
__device__ __forceinline__ void fft_frames(float2 (&x)[NB][16], float2* const (&sm)[NB], const float2* __restrict__ tw, int t, int f,
                                           Hook after_first_sync = Hook()) {
  fft_passes_from<N, 0, INV, NB, TAIL0, Hook>(x, sm, tw, t, f, after_first_sync);
}

// single-frame convenience wrapper
template <int N, bool INV>
__device__ __forceinline__ void fft_frame(float2 (&x)[16], float2* __restrict__ smem, const float2* __restrict__ tw, int t, int f) {
  float2* const sm[1] = {smem};
  float2 y[1][16];
#pragma unroll
  for (int m = 0; m < 16; ++m) y[0][m] = x[m];
  fft_frames<N, INV, 1>(y, sm, tw, t, f);
#pragma unroll
  for (int m = 0; m < 16; ++m) x[m] = y[0][m];
}

}  // namespace ae
