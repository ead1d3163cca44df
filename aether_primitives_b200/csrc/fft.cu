// fft.cu — K2 batched power-of-two FFT and K2b any-length fallback (sm_100a).
// Replaces rustfft under fft::Cfft (src/fft.rs:134-235).  Scale::{SN,N,X} (src/fft.rs:22-37) is
// folded into the last pass as one separately rounded multiply, which reproduces the
// reference's "transform, then vec_scale" rounding sequence on this kernel's own output.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "async_copy.cuh"
#include "fft_device.cuh"
#include "internal.h"

namespace ae {

// -------------------------------------------------------------------------------------------------
// power-of-two kernel: F frames per CTA, T = N/16 threads per frame, algorithmic traffic
// 16 B/sample (8 read + 8 written), 5*N*log2(N) flop per frame.
// -------------------------------------------------------------------------------------------------
template <int N>
struct FftLaunch {
  static constexpr int T = FftCfg<N>::T;
  // frames per CTA: 256-thread CTAs (named barriers allow <= 15 frame slots when T >= 32)
  static constexpr int F = T >= 256 ? 1 : (256 / T);
  static constexpr int THREADS = F * T;
  static constexpr int MINB = THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1);  // <= 64 registers at 256 threads
  static constexpr size_t smem(bool staged) {
    return (size_t)F * (FftCfg<N>::SMEM_ELEMS + (staged ? N : 0)) * sizeof(float2) + (staged ? F * sizeof(uint64_t) : 0);
  }
};

// STAGED: each frame slot prefetches its NEXT frame with one cp.async.bulk (TMA) into a staging buffer
// while it transforms the current one (mbarrier completion), so twice the bytes are in flight per SM
// without spending registers; needs a 16-byte aligned input.
template <int N, bool INV, bool STAGED>
__global__ void __launch_bounds__(FftLaunch<N>::THREADS, FftLaunch<N>::MINB)
fft_pow2_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, size_t frames,
                float scale, int do_scale) {
  using C = FftCfg<N>;
  using LC = FftLaunch<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  const int f = threadIdx.x / C::T;
  const int t = threadIdx.x % C::T;
  float2* sm = smem + (size_t)f * (C::SMEM_ELEMS + (STAGED ? N : 0));
  float2* xin = sm + C::SMEM_ELEMS;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)LC::F * (C::SMEM_ELEMS + (STAGED ? N : 0))) + f;
  const size_t stride = (size_t)gridDim.x * LC::F;
  size_t frame = (size_t)blockIdx.x * LC::F + f;
  const bool wide = (((uintptr_t)in | (uintptr_t)out) % 32) == 0;   // 256-bit accesses allowed (used when T == 1)
  if (STAGED) {
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
      if (frame < frames) {
        mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
        bulk_g2s(xin, in + frame * N, N * (uint32_t)sizeof(float2), bar);
      }
    }
    __syncthreads();
  }
  uint32_t phase = 0;
  // persistent: each frame slot (C::T threads) walks its own frames and synchronises only with
  // itself, so the slots resident on an SM drift apart and overlap their load / compute / store phases
  for (; frame < frames; frame += stride) {
    float2 x[1][16];
    if (STAGED) {
      mbar_wait(bar, phase);
      phase ^= 1u;
#pragma unroll
      for (int m = 0; m < 16; ++m) x[0][m] = xin[t + m * C::T];
    } else {
      const float2* src = in + frame * N;
      if (C::T == 1 && wide) {
        // N = 16: the thread owns the whole 128-byte frame - four 256-bit loads instead of sixteen 64-bit
        // loads that each touch 32 different lines per warp
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 a, b;
          ld_stream_256(src + 4 * j, a, b);
          x[0][4 * j] = make_float2(a.x, a.y); x[0][4 * j + 1] = make_float2(a.z, a.w);
          x[0][4 * j + 2] = make_float2(b.x, b.y); x[0][4 * j + 3] = make_float2(b.z, b.w);
        }
      } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) x[0][m] = ld_stream(src + t + m * C::T);
      }
    }
    auto prefetch = [&]() {
      if (STAGED && t == 0 && frame + stride < frames) {
        mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
        bulk_g2s(xin, in + (frame + stride) * N, N * (uint32_t)sizeof(float2), bar);
      }
    };
    float2* const smv[1] = {sm};
    if (C::NP > 1) {
      fft_frames<N, INV, 1, false>(x, smv, tw, t, f, prefetch);
    } else {  // single pass: no slot barrier inside, make one so the staging buffer can be refilled
      fft_frames<N, INV, 1, false>(x, smv, tw, t, f);
      if (STAGED) { frame_sync<C::T>(f); prefetch(); }
    }
    float2* dst = out + frame * N;
    if (do_scale) {
#pragma unroll
      for (int m = 0; m < 16; ++m) x[0][m] = cx_scale_exact(x[0][m], scale);
    }
    if (C::T == 1 && wide) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        st_stream_256(dst + 4 * j, make_float4(x[0][4 * j].x, x[0][4 * j].y, x[0][4 * j + 1].x, x[0][4 * j + 1].y),
                      make_float4(x[0][4 * j + 2].x, x[0][4 * j + 2].y, x[0][4 * j + 3].x, x[0][4 * j + 3].y));
    } else {
#pragma unroll
      for (int m = 0; m < 16; ++m) st_stream(dst + t + m * C::T, x[0][m]);
    }
    // the next frame's first pass stores into the same shared frame: everyone must be past its reads
    if (C::NP > 1) frame_sync<C::T>(f);
  }
}

template <int N>
static void launch_pow2_n(const float2* in, float2* out, size_t frames, const float2* tw, bool inverse, bool do_scale,
                          float scale, cudaStream_t st) {
  using LC = FftLaunch<N>;
  const size_t want = (frames + LC::F - 1) / LC::F;
  // Measured on B200 (tools/fft_quick.py): with 4 CTAs/SM of plain coalesced loads the kernel already
  // sits at 93-95 % of the measured HBM peak; the TMA-staged variant is 1-3 % slower because its
  // staging buffers cut the resident CTAs from 4 to 3.  It stays selectable (AE_FFT_TMA=1) and tested.
  static const char* use_tma = getenv("AE_FFT_TMA");
  const bool staged = ((uintptr_t)in % 16) == 0 && use_tma && N >= 64 && N <= 2048;
  auto launch = [&](auto kern, bool stg) {
    const size_t smem = LC::smem(stg);
    const size_t resident = resident_ctas((const void*)kern, LC::THREADS, smem);  // grid = SM count x resident CTAs
    const unsigned grid = (unsigned)(want < resident ? want : resident);
    kern<<<grid, LC::THREADS, smem, st>>>(in, out, tw, frames, scale, do_scale);
  };
  if (inverse) { if (staged) launch(fft_pow2_kernel<N, true, true>, true); else launch(fft_pow2_kernel<N, true, false>, false); }
  else { if (staged) launch(fft_pow2_kernel<N, false, true>, true); else launch(fft_pow2_kernel<N, false, false>, false); }
}

bool fft_pow2_supported(size_t n) { return n >= 16 && n <= 16384 && (n & (n - 1)) == 0; }

// Per-thread twiddle table of the power-of-two kernels (layout described in fft_device.cuh):
// for every pass p >= 1 and thread t the handful of exp(-2 pi i e/n) values that thread needs,
// evaluated in f64 and rounded to f32.
template <int N>
static void thread_twiddles_n(std::vector<float2>& out) {
  using C = FftCfg<N>;
  out.assign((size_t)(C::TW_ROWS > 0 ? C::TW_ROWS : 1) * C::T, make_float2(1.f, 0.f));
  auto W = [](long long e) {
    const double a = -2.0 * 3.14159265358979323846 * (double)(e % N) / (double)N;
    return make_float2((float)std::cos(a), (float)std::sin(a));
  };
  for (int p = 1; p < C::NP; ++p) {
    const int R = C::radix(p), NS = C::ns(p), row0 = C::tw_row0(p);
    for (int t = 0; t < C::T; ++t) {
      if (R == 16) {
        const long long u = (long long)(t & (NS - 1)) * (N / (NS * 16));
        const int mult[6] = {1, 2, 3, 4, 8, 12};
        for (int e = 0; e < 6; ++e) out[(size_t)(row0 + e) * C::T + t] = W(mult[e] * u);
      } else {
        for (int r = 1; r < R; ++r) out[(size_t)(row0 + r - 1) * C::T + t] = W((long long)r * t);
      }
    }
  }
}
void fft_thread_twiddles(size_t n, std::vector<float2>& out) {
  switch (n) {
#define AE_CASE(NN) case NN: thread_twiddles_n<NN>(out); break;
    AE_CASE(16) AE_CASE(32) AE_CASE(64) AE_CASE(128) AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048)
    AE_CASE(4096) AE_CASE(8192) AE_CASE(16384)
#undef AE_CASE
    default: out.clear();
  }
}

void launch_fft_pow2(const float2* in, float2* out, size_t n, size_t frames, const float2* tw, bool inverse, bool do_scale,
                     float scale, cudaStream_t st) {
  if (frames == 0) return;
  switch (n) {
#define AE_CASE(NN) case NN: launch_pow2_n<NN>(in, out, frames, tw, inverse, do_scale, scale, st); break;
    AE_CASE(16) AE_CASE(32) AE_CASE(64) AE_CASE(128) AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048)
    AE_CASE(4096) AE_CASE(8192) AE_CASE(16384)
#undef AE_CASE
    default: note_unsupported_launch("power-of-two FFT kernel: length must be 16..16384");
  }
}

// -------------------------------------------------------------------------------------------------
// Large power-of-two lengths, 2^15 <= N <= 2^24: four-step FFT, N = N1*N2, as TWO passes of a
// column-FFT kernel over the frame viewed as a matrix (32 B/sample of traffic instead of one
// global pass per radix-4 stage):
//   pass 1: x[n1][n2] -> column FFTs over n1 (length N1, stride N2), times W_N^(n2*k1), written
//           TRANSPOSED to scratch[n2][k1]
//   pass 2: scratch[n2][k1] -> column FFTs over n2 (length N2, stride N1), in place by position, so the
//           result lands at out[k2*N1 + k1] = X[k1 + N1*k2]: natural order, no further transpose.
// A CTA owns CT adjacent columns; thread = (t, c) with c fastest, so every global access is a run of
// CT contiguous cf32 per row.  W_N^e is formed from two 4096-entry tables, W^(e>>12 <<12) * W^(e&4095).
// -------------------------------------------------------------------------------------------------
constexpr int col_default_ct(int n1) { return n1 >= 4096 ? 4 : (n1 >= 1024 ? 8 : 16); }
template <int N1, int CTV = col_default_ct(N1)>
struct ColLaunch {
  static constexpr int T = FftCfg<N1>::T;
  static constexpr int CT = CTV;
  static constexpr int THREADS = CT * T;
  // lanes of a warp are different COLUMNS at the same offset: skew each column's buffer by one cf32 so
  // they fall into different banks (SMEM_ELEMS is a multiple of 16 cf32 = one 128-byte bank row)
  static constexpr int CSTRIDE = FftCfg<N1>::SMEM_ELEMS + 1;
  static constexpr size_t SMEM = (size_t)CT * CSTRIDE * sizeof(float2);
};

template <int N1, bool INV, bool FIRST, int CTV = col_default_ct(N1)>
__global__ void __launch_bounds__(ColLaunch<N1, CTV>::THREADS, 1024 / ColLaunch<N1, CTV>::THREADS)
fft_col_kernel(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ tw, const float2* __restrict__ wlo,
               const float2* __restrict__ whi, size_t stride, size_t frame_elems, float scale, int do_scale) {
  using C = FftCfg<N1>;
  using LC = ColLaunch<N1, CTV>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  const int c = threadIdx.x % LC::CT;
  const int t = threadIdx.x / LC::CT;
  const size_t col = (size_t)blockIdx.x * LC::CT + c;
  const float2* src = in + (size_t)blockIdx.y * frame_elems + col;
  float2 x[1][16];
#pragma unroll
  for (int m = 0; m < 16; ++m) x[0][m] = ld_stream(src + (size_t)(t + m * C::T) * stride);
  float2* const smv[1] = {smem + (size_t)c * LC::CSTRIDE};
  fft_frames<N1, INV, 1, false>(x, smv, tw, t, -1);
  if (FIRST) {
    // twiddle W_N^(n2*k1), n2 = col, k1 = t + m*T, then transpose through shared memory
    __syncthreads();  // every thread is past its last read of the FFT buffers
    // exponent col*(t + m*T) = e0 + (4a + b)*es: W^(e0 + 4a es) and W^(b es) come from the two-level table
    // (7 look-up pairs), the sixteen factors are their products - one product deeper than a look-up per
    // element (which made the pass L1-bound: 32 scattered reads per thread), still inside the accuracy
    // bar (error vs f64 truth within 3 dB of the oracle's); deriving all sixteen from two look-ups was not.
    auto wn = [&](unsigned e) { return cx_mul(__ldg(whi + (e >> 12)), __ldg(wlo + (e & 4095u))); };   // e < N <= 2^24
    const unsigned e0 = (unsigned)col * (unsigned)t, es = (unsigned)col * (unsigned)C::T;
    const float2 wa[4] = {wn(e0), wn(e0 + 4u * es), wn(e0 + 8u * es), wn(e0 + 12u * es)};
    const float2 wb[4] = {make_float2(1.0f, 0.0f), wn(es), wn(2u * es), wn(3u * es)};
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const float2 w = (m & 3) ? cx_mul(wa[m >> 2], wb[m & 3]) : wa[m >> 2];
      smv[0][fft_pad(t + m * C::T)] = INV ? cx_mul_conj(x[0][m], w) : cx_mul(x[0][m], w);
    }
    __syncthreads();
    // tile (c, k1) -> scratch[(col0 + c) * N1 + k1]: rows of N1 contiguous cf32
    float2* dst = out + (size_t)blockIdx.y * frame_elems + (size_t)blockIdx.x * LC::CT * N1;
    for (int i = threadIdx.x; i < LC::CT * N1; i += LC::THREADS) {
      const int cc = i / N1, k = i % N1;
      st_stream(dst + (size_t)cc * N1 + k, smem[(size_t)cc * LC::CSTRIDE + fft_pad(k)]);
    }
  } else {
    float2* dst = out + (size_t)blockIdx.y * frame_elems + col;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      float2 v = x[0][m];
      if (do_scale) v = cx_scale_exact(v, scale);
      st_stream(dst + (size_t)(t + m * C::T) * stride, v);
    }
  }
}

template <int N1, bool FIRST, int CTV = col_default_ct(N1)>
static void launch_col_n(const float2* in, float2* out, const float2* tw, const float2* wlo, const float2* whi, size_t ncols,
                         size_t frames, size_t frame_elems, bool inverse, bool do_scale, float scale, cudaStream_t st) {
  using LC = ColLaunch<N1, CTV>;
  auto launch = [&](auto kern) {
    (void)resident_ctas((const void*)kern, LC::THREADS, LC::SMEM);   // shared-memory opt-in, once per device
    for (size_t f0 = 0; f0 < frames; f0 += 32768) {
      const size_t fc = frames - f0 < 32768 ? frames - f0 : 32768;
      const dim3 grid((unsigned)(ncols / LC::CT), (unsigned)fc);
      kern<<<grid, LC::THREADS, LC::SMEM, st>>>(in + f0 * frame_elems, out + f0 * frame_elems, tw, wlo, whi, ncols, frame_elems, scale,
                                               do_scale);
    }
  };
  if (inverse) launch(fft_col_kernel<N1, true, FIRST, CTV>);
  else launch(fft_col_kernel<N1, false, FIRST, CTV>);
}

bool fft_big_supported(size_t n) { return n > 16384 && n <= ((size_t)1 << 24) && (n & (n - 1)) == 0; }
void fft_big_split(size_t n, size_t* n1, size_t* n2) {
  int l = 0;
  while (((size_t)1 << l) < n) ++l;
  *n1 = (size_t)1 << ((l + 1) / 2);
  *n2 = (size_t)1 << (l / 2);
}

// in -> scratch (pass 1) -> out (pass 2); scratch holds n*frames cf32; in may equal out
void launch_fft_big(const float2* in, float2* out, float2* scratch, size_t n, size_t frames, const float2* tw1, const float2* tw2,
                    const float2* wlo, const float2* whi, bool inverse, bool do_scale, float scale, cudaStream_t st) {
  if (frames == 0) return;
  size_t n1, n2;
  fft_big_split(n, &n1, &n2);
  switch (n1) {
#define AE_CASE(NN) case NN: launch_col_n<NN, true>(in, scratch, tw1, wlo, whi, n2, frames, n, inverse, false, 1.0f, st); break;
    AE_CASE(128) AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096)
#undef AE_CASE
    default: return;
  }
  switch (n2) {
#define AE_CASE(NN) case NN: launch_col_n<NN, false>(scratch, out, tw2, wlo, whi, n1, frames, n, inverse, do_scale, scale, st); break;
    AE_CASE(128) AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096)
#undef AE_CASE
    default: return;
  }
}

// -------------------------------------------------------------------------------------------------
// K2b any-length fallback (Cfft::with_len accepts any len; N = 100 is a reference test,
// src/vecops.rs:445-463).  One global-memory Stockham pass per factor p of N; each thread produces
// ONE output  out[(j-k)*p + k + q*NS] = sum_r in[j + r*N/p] * W_N^( r*k*N/(NS*p) + ((r*q) mod p)*N/p ).
// Correctness-first: O(N * sum(p)) work, f32 accumulation.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fft_generic_pass_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                               const float2* __restrict__ tw, unsigned n, unsigned p, unsigned ns,
                                                               int inverse, int do_scale, float scale) {
  const unsigned o = blockIdx.x * 256 + threadIdx.x;  // output slot: (j, q)
  if (o >= n) return;
  const size_t frame = blockIdx.y;
  const unsigned m = n / p;       // butterflies
  const unsigned j = o % m, q = o / m;
  const unsigned k = j % ns;
  const unsigned tws = n / (ns * p), wp = n / p;
  const float2* src = in + frame * n;
  float2 acc = make_float2(0.0f, 0.0f);
  for (unsigned r = 0; r < p; ++r) {
    const unsigned long long e = (unsigned long long)r * k * tws + (unsigned long long)((r * (unsigned long long)q) % p) * wp;
    float2 w = __ldg(tw + (unsigned)(e % n));
    if (inverse) w.y = -w.y;
    cx_fma(acc, src[j + r * m], w);
  }
  if (do_scale) acc = cx_scale_exact(acc, scale);
  out[frame * n + (size_t)(j - k) * p + k + (size_t)q * ns] = acc;
}

// Any length up to 6144 in ONE kernel: the frames of a CTA live in shared memory (two ping-pong buffers,
// padded by one cf32 per 32), one Stockham pass per factor.  The first pass reads global memory, the
// last one writes it.  Radices 2/3/4/5/7/8/11/13 run as register butterflies (one butterfly = one
// thread: P loads, P-1 table twiddles, P stores); any other (prime) factor falls back to one output per
// thread with P terms.  Several frames share a CTA when the length is small (N = 100, the reference's
// own test length: 20 frames per CTA), so all 256 threads have work in every pass.  The inverse
// transform is conj -> forward -> conj, applied at the first load and the last store.
constexpr int kGenSmemMaxN = 6144;
constexpr int kGenMaxRadices = 16;
// per pass: radix p, m = n/p butterflies, ns = product of the earlier radices, and 2^32/d "magic"
// reciprocals so the index arithmetic needs no integer division (exact while x*d < 2^32)
struct GenPass { unsigned p, m, ns, tws, mg_m, mg_ns; };
struct GenRadices { GenPass pass[kGenMaxRadices]; unsigned mg_n; int count; };
// floor(x / d) with magic = ceil(2^32 / d) (0 encodes d == 1).  mul.hi through the intrinsic: the same
// arithmetic written as a 64-bit product and shift was folded by nvcc 12.9 into x*hi32(magic*-d)+x,
// which is wrong (N = 12 crashed with an illegal address).
__device__ __forceinline__ unsigned gen_div(unsigned x, unsigned magic) { return magic ? __umulhi(x, magic) : x; }

__device__ __forceinline__ unsigned gen_pad(unsigned i) { return i + (i >> 5); }

// forward DFT of odd prime length P on registers.  root[r] = exp(-2 pi i r / P), r = 1..(P-1)/2:
//   X[q], X[P-q] = x0 + sum_r cos(2 pi r q / P) (x_r + x_{P-r})  -/+  i sum_r sin(2 pi r q / P) (x_r - x_{P-r})
// The cosine sum is evaluated in the DC-exact form: with a_r = x_r + x_{P-r} and sum_r cos(2 pi r q/P) = -1/2,
//   x0 + sum_r c_rq a_r = (x0 - a_1/2) + sum_{r>=2} c_rq (a_r - a_1),
// so a CONSTANT input gives exactly zero in every bin but the first, as radix-2/4/8 butterflies do.  The
// reference's own FFT tests (src/vecops.rs:445-463: N = 100, constant input, SN round trip, assert_evm! at -80,
// i.e. equality to the bit) rely on exactly that; the plain sum leaves ~1e-7 in the second radix-5 pass.
template <int P>
__device__ __forceinline__ void odd_dft(float2 (&v)[P], const float2 (&root)[(P - 1) / 2 + 1]) {
  constexpr int H = (P - 1) / 2;
  float2 a[H + 1], b[H + 1];
#pragma unroll
  for (int r = 1; r <= H; ++r) { a[r] = cx_add(v[r], v[P - r]); b[r] = cx_sub(v[r], v[P - r]); }
  float2 y0 = v[0];
#pragma unroll
  for (int r = 1; r <= H; ++r) y0 = cx_add(y0, a[r]);
  const float2 base = make_float2(fmaf(-0.5f, a[1].x, v[0].x), fmaf(-0.5f, a[1].y, v[0].y));   // x0 - a_1/2 (exact halving)
  float2 da[H + 1];
#pragma unroll
  for (int r = 2; r <= H; ++r) da[r] = cx_sub(a[r], a[1]);
  float2 o[P];
  o[0] = y0;
#pragma unroll
  for (int q = 1; q <= H; ++q) {
    float2 e = base, f = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int r = 1; r <= H; ++r) {
      const int idx = (r * q) % P;                       // compile-time after unrolling
      const float c = idx <= H ? root[idx].x : root[P - idx].x;
      const float sn = idx <= H ? -root[idx].y : root[P - idx].y;   // sin(2 pi idx / P)
      if (r >= 2) { e.x = fmaf(c, da[r].x, e.x); e.y = fmaf(c, da[r].y, e.y); }
      f.x = fmaf(sn, b[r].x, f.x); f.y = fmaf(sn, b[r].y, f.y);
    }
    o[q] = make_float2(e.x + f.y, e.y - f.x);            // e - i f
    o[P - q] = make_float2(e.x - f.y, e.y + f.x);        // e + i f
  }
#pragma unroll
  for (int i = 0; i < P; ++i) v[i] = o[i];
}

// Composite radix P = P1 * P2 with P1 in {2, 4}, P2 an odd prime, coprime: Good-Thomas prime-factor mapping, no
// twiddles inside the butterfly.   n = (P2 n1 + P1 n2) mod P,  k = (P2 c1 k1 + P1 c2 k2) mod P  with
// c1 = P2^-1 mod P1, c2 = P1^-1 mod P2:   X[k] = sum_n1 W_P1^(n1 k1) sum_n2 W_P2^(n2 k2) x[n].
// One pass of radix 10 replaces a radix-5 and a radix-2 pass (N = 1000: three passes instead of four; the passes are
// bound by the shared-memory pipe).  A constant input still gives exact zeros off DC (odd_dft is DC-exact, the
// power-of-two butterflies are sums and differences).
constexpr int pfa_inv_mod(int a, int m) {
  for (int x = 1; x < m; ++x)
    if ((a * x) % m == 1) return x;
  return 1;
}
template <int P1, int P2>
__device__ __forceinline__ void pfa_dft(float2 (&v)[P1 * P2], const float2 (&root)[(P2 - 1) / 2 + 1]) {
  constexpr int P = P1 * P2;
  float2 s[P1][P2];
#pragma unroll
  for (int n1 = 0; n1 < P1; ++n1) {
#pragma unroll
    for (int n2 = 0; n2 < P2; ++n2) s[n1][n2] = v[(P2 * n1 + P1 * n2) % P];
    odd_dft<P2>(s[n1], root);
  }
  constexpr int c1 = pfa_inv_mod(P2 % P1, P1), c2 = pfa_inv_mod(P1 % P2, P2);
#pragma unroll
  for (int k2 = 0; k2 < P2; ++k2) {
    float2 c[P1];
#pragma unroll
    for (int n1 = 0; n1 < P1; ++n1) c[n1] = s[n1][k2];
    Dft<P1, false>::run(c);
#pragma unroll
    for (int k1 = 0; k1 < P1; ++k1) v[(P2 * c1 * k1 + P1 * c2 * k2) % P] = c[k1];
  }
}
// the odd prime inside a radix handled by gen_butterfly_pass (0: a power of two)
__host__ __device__ constexpr int gen_odd_part(int p) { return (p & 1) ? p : (p == 6 || p == 12) ? 3 : (p == 10) ? 5 : 0; }

template <int P, bool SRC_SMEM, bool DST_SMEM>
__device__ __forceinline__ void gen_butterfly_pass(const float2* src, float2* dst, const float2* __restrict__ tw, unsigned n, const GenPass& gp,
                                                   unsigned slots, bool conj_in, bool conj_out, bool do_scale, float scale) {
  const unsigned m = gp.m, tws = gp.tws, ns = gp.ns;
  constexpr int PO = gen_odd_part(P);                    // roots of unity of the odd factor: exp(-2 pi i r / PO) = tw[r n / PO]
  constexpr int H = PO ? (PO - 1) / 2 : 0;
  float2 root[H + 1];
  if constexpr (PO != 0) {
#pragma unroll
    for (int r = 1; r <= H; ++r) root[r] = __ldg(tw + r * (m * (P / PO)));
  }
  for (unsigned item = threadIdx.x; item < slots * m; item += blockDim.x) {
        const unsigned s = gen_div(item, gp.mg_m), j = item - s * m, k = j - gen_div(j, gp.mg_ns) * ns, sb = s * n;
    float2 v[P];
#pragma unroll
    for (int r = 0; r < P; ++r) {
      const unsigned idx = sb + j + r * m;
      v[r] = SRC_SMEM ? src[gen_pad(idx)] : ld_stream(src + idx);
      if (conj_in) v[r].y = -v[r].y;
    }
    if (ns > 1) {
      // W^(r k tws), r = 1..P-1: one table read, the higher powers as balanced products (the pass is bound
      // by the L1/shared-memory pipe, not by FP32 issue; each product costs about one ulp)
      float2 w[P];
      w[1] = __ldg(tw + k * tws);
      if constexpr (P >= 10) {
        // the composite radices reach W^9 .. W^11: three table reads (r k tws < n, no wrap) keep every power within two
        // products of a table entry, so the pass is as accurate as the prime-radix passes it replaces
        w[2] = __ldg(tw + 2 * k * tws);
        w[4] = __ldg(tw + 4 * k * tws);
        w[3] = cx_mul(w[1], w[2]);
#pragma unroll
        for (int r = 5; r < P; ++r) w[r] = cx_mul(w[r - 4], w[4]);
      } else {
#pragma unroll
        for (int r = 2; r < P; ++r) w[r] = cx_mul(w[r / 2], w[r - r / 2]);
      }
#pragma unroll
      for (int r = 1; r < P; ++r) v[r] = cx_mul(v[r], w[r]);
    }
    if constexpr (P & 1) odd_dft<P>(v, root);
    else if constexpr (PO != 0) pfa_dft<P / PO, PO>(v, root);
    else Dft<P, false>::run(v);
    const unsigned d0 = sb + (j - k) * P + k;
#pragma unroll
    for (int q = 0; q < P; ++q) {
      float2 y = v[q];
      const unsigned idx = d0 + q * ns;
      if (DST_SMEM) {
        dst[gen_pad(idx)] = y;
      } else {
        if (conj_out) y.y = -y.y;
        if (do_scale) y = cx_scale_exact(y, scale);
        st_stream(dst + idx, y);
      }
    }
  }
}

// any radix p: one output per thread, p terms (exponent of W_n for term r: r*k*tws + ((r*q) mod p)*n/p)
template <bool SRC_SMEM, bool DST_SMEM>
__device__ __forceinline__ void gen_output_pass(const float2* src, float2* dst, const float2* __restrict__ tw, unsigned n, const GenPass& gp,
                                                unsigned mg_n, unsigned slots, bool conj_in, bool conj_out, bool do_scale, float scale) {
  const unsigned p = gp.p, m = gp.m, tws = gp.tws, wp = gp.m, ns = gp.ns;
  for (unsigned o = threadIdx.x; o < slots * n; o += blockDim.x) {
    const unsigned s = gen_div(o, mg_n), oo = o - s * n, q = gen_div(oo, gp.mg_m), j = oo - q * m, k = j - gen_div(j, gp.mg_ns) * ns, sb = s * n;
    float2 acc = make_float2(0.0f, 0.0f);
    const unsigned step1 = k * tws;
    unsigned e1 = 0, rq = 0;
    for (unsigned r = 0; r < p; ++r) {
      unsigned e = e1 + rq * wp;
      if (e >= n) e -= n;
      const float2 w = __ldg(tw + e);
      const unsigned idx = sb + j + r * m;
      float2 x = SRC_SMEM ? src[gen_pad(idx)] : src[idx];
      if (conj_in) x.y = -x.y;
      cx_fma(acc, x, w);
      e1 += step1;
      rq += q;
      if (rq >= p) rq -= p;
    }
    const unsigned idx = sb + (j - k) * p + k + q * ns;
    if (DST_SMEM) {
      dst[gen_pad(idx)] = acc;
    } else {
      if (conj_out) acc.y = -acc.y;
      if (do_scale) acc = cx_scale_exact(acc, scale);
      dst[idx] = acc;
    }
  }
}

__host__ __device__ constexpr bool gen_has_butterfly(unsigned p) {
  return p == 2 || p == 3 || p == 4 || p == 5 || p == 6 || p == 7 || p == 8 || p == 10 || p == 11 || p == 12 || p == 13;
}
template <bool SRC_SMEM, bool DST_SMEM>
__device__ __forceinline__ void gen_pass(const GenPass& gp, unsigned mg_n, const float2* src, float2* dst, const float2* __restrict__ tw,
                                         unsigned n, unsigned slots, bool conj_in, bool conj_out, bool do_scale, float scale) {
  switch (gp.p) {
#define AE_P(PP) case PP: gen_butterfly_pass<PP, SRC_SMEM, DST_SMEM>(src, dst, tw, n, gp, slots, conj_in, conj_out, do_scale, scale); break;
    AE_P(2) AE_P(3) AE_P(4) AE_P(5) AE_P(6) AE_P(7) AE_P(8) AE_P(10) AE_P(11) AE_P(12) AE_P(13)
#undef AE_P
    default: gen_output_pass<SRC_SMEM, DST_SMEM>(src, dst, tw, n, gp, mg_n, slots, conj_in, conj_out, do_scale, scale); break;
  }
}

#ifndef AE_GEN_MINB
#define AE_GEN_MINB 3
#endif
#ifndef AE_GEN_ELEMS
#define AE_GEN_ELEMS 4096
#endif
__global__ void __launch_bounds__(256, AE_GEN_MINB) fft_generic_smem_kernel(const float2* __restrict__ in, float2* __restrict__ out,
                                                               const float2* __restrict__ tw, unsigned n, size_t frames, unsigned slots,
                                                               const __grid_constant__ GenRadices rad, int inverse, int do_scale,
                                                               float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf0 = reinterpret_cast<float2*>(smem_raw);
  float2* buf1 = buf0 + gen_pad(slots * n) + 1;
  // the per-pass plan is read with a run-time index: keep it in shared memory
  __shared__ GenPass plan[kGenMaxRadices];
  if (threadIdx.x < rad.count) plan[threadIdx.x] = rad.pass[threadIdx.x];
  __syncthreads();
  for (size_t frame0 = (size_t)blockIdx.x * slots; frame0 < frames; frame0 += (size_t)gridDim.x * slots) {
    const unsigned here = frames - frame0 < slots ? (unsigned)(frames - frame0) : slots;
    const float2* gin = in + frame0 * n;
    float2* gout = out + frame0 * n;
    float2* cur = buf0;
    float2* nxt = buf1;
    // a first factor without a register butterfly reads every input from many threads: stage the frames
    // in shared memory first (also keeps a single-pass in-place transform free of read/write races)
    const bool staged = !gen_has_butterfly(plan[0].p);
    if (staged) {
      for (unsigned i = threadIdx.x; i < here * n; i += blockDim.x) nxt[gen_pad(i)] = ld_stream(gin + i);
      __syncthreads();
    }
    for (int pi = 0; pi < rad.count; ++pi) {
      const GenPass gp = plan[pi];
      const bool first = pi == 0, last = pi == rad.count - 1;
      if (first && staged) {
        if (last) gen_pass<true, false>(gp, rad.mg_n, nxt, gout, tw, n, here, inverse, inverse, do_scale, scale);
        else gen_pass<true, true>(gp, rad.mg_n, nxt, cur, tw, n, here, inverse, false, false, scale);
      }
      else if (first && last) gen_pass<false, false>(gp, rad.mg_n, gin, gout, tw, n, here, inverse, inverse, do_scale, scale);
      else if (first) gen_pass<false, true>(gp, rad.mg_n, gin, cur, tw, n, here, inverse, false, false, scale);
      else if (last) gen_pass<true, false>(gp, rad.mg_n, cur, gout, tw, n, here, false, inverse, do_scale, scale);
      else gen_pass<true, true>(gp, rad.mg_n, cur, nxt, tw, n, here, false, false, false, scale);
      __syncthreads();
      if (!first) { float2* t = cur; cur = nxt; nxt = t; }
    }
  }
}

static unsigned gen_slots(size_t n) {
  size_t s = AE_GEN_ELEMS / n;
  return (unsigned)(s < 1 ? 1 : (s > 64 ? 64 : s));
}

void launch_fft_generic(const float2* in, float2* out, float2* scratch, size_t n, size_t frames, const float2* tw,
                        const uint32_t* radices, int n_radices, bool inverse, bool do_scale, float scale, cudaStream_t st) {
  if (frames == 0 || n == 0) return;
  if (n <= (size_t)kGenSmemMaxN && n_radices >= 1 && n_radices <= kGenMaxRadices) {
    // odd radices first (their stride-P stores are conflict-free), then 5 x 2 -> 10, 3 x 4 -> 12, 3 x 2 -> 6 (fewer
    // passes: each one moves the whole frame through shared memory), then the powers of two merged into 8s
    static const bool composite = getenv("AE_FFT_NO_COMPOSITE") == nullptr;
    unsigned order[kGenMaxRadices];
    int cnt = 0;
    unsigned twos = 0, threes = 0, fives = 0;
    for (int i = 0; i < n_radices; ++i) {
      if (radices[i] == 2) twos += 1;
      else if (radices[i] == 4) twos += 2;
      else if (radices[i] == 8) twos += 3;
      else if (radices[i] == 3 && composite) threes += 1;
      else if (radices[i] == 5 && composite) fives += 1;
      else order[cnt++] = radices[i];
    }
    unsigned tens = 0, twelves = 0, sixes = 0;
    while (fives && twos) { ++tens; --fives; --twos; }
    while (threes && twos >= 2) { ++twelves; --threes; twos -= 2; }
    while (threes && twos) { ++sixes; --threes; --twos; }
    for (; threes; --threes) order[cnt++] = 3;
    for (; fives; --fives) order[cnt++] = 5;
    for (; tens; --tens) order[cnt++] = 10;
    for (; twelves; --twelves) order[cnt++] = 12;
    for (; sixes; --sixes) order[cnt++] = 6;
    for (; twos >= 3 && cnt < kGenMaxRadices; twos -= 3) order[cnt++] = 8;
    if (twos == 2) order[cnt++] = 4;
    if (twos == 1) order[cnt++] = 2;
    auto magic = [](unsigned d) { return d <= 1 ? 0u : (unsigned)((0x100000000ull + d - 1) / d); };
    GenRadices rad;
    rad.count = cnt;
    rad.mg_n = magic((unsigned)n);
    unsigned ns = 1;
    for (int i = 0; i < cnt; ++i) {
      GenPass& gp = rad.pass[i];
      gp.p = order[i];
      gp.m = (unsigned)n / gp.p;
      gp.ns = ns;
      gp.tws = (unsigned)n / (ns * gp.p);
      gp.mg_m = magic(gp.m);
      gp.mg_ns = magic(ns);
      ns *= gp.p;
    }
    const unsigned slots = gen_slots(n);
    const size_t smem = 2 * ((size_t)slots * n + (slots * n) / 32 + 2) * sizeof(float2);
    const size_t resident = resident_ctas((const void*)fft_generic_smem_kernel, 256, smem);
    const size_t want = (frames + slots - 1) / slots;
    fft_generic_smem_kernel<<<(unsigned)(want < resident ? want : resident), 256, smem, st>>>(in, out, tw, (unsigned)n, frames, slots, rad,
                                                                                          inverse, do_scale, scale);
    return;
  }
  // scratch holds 2*n*frames cf32: ping-pong halves; the last pass writes `out`
  constexpr size_t kMaxY = 32768;  // gridDim.y limit is 65535
  for (size_t f0 = 0; f0 < frames; f0 += kMaxY) {
    const size_t fc = frames - f0 < kMaxY ? frames - f0 : kMaxY;
    float2* bufs[2] = {scratch + f0 * n, scratch + n * frames + f0 * n};
    const float2* cur = in + f0 * n;
    float2* o = out + f0 * n;
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)fc);
    if (n == 1 || n_radices <= 1) {
      // single pass: never read and write the same buffer
      if (cur == o) {
        cudaMemcpyAsync(bufs[0], cur, fc * n * sizeof(float2), cudaMemcpyDeviceToDevice, st);
        cur = bufs[0];
      }
      fft_generic_pass_kernel<<<grid, 256, 0, st>>>(cur, o, tw, (unsigned)n, (unsigned)n, 1, inverse, do_scale, scale);
      continue;
    }
    unsigned ns = 1;
    for (int i = 0; i < n_radices; ++i) {
      const bool last = (i == n_radices - 1);
      float2* dst = last ? o : bufs[i & 1];
      fft_generic_pass_kernel<<<grid, 256, 0, st>>>(cur, dst, tw, (unsigned)n, radices[i], ns, inverse, last && do_scale, scale);
      cur = dst;
      ns *= radices[i];
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Bluestein (chirp-z) for lengths the mixed-radix kernels handle badly: non-power-of-two N > 6144 and
// N with a large prime factor.  With w[n] = exp(-i pi n^2 / N):
//     X[k] = w[k] * sum_n (x[n] w[n]) * conj(w[k-n])
// i.e. one circular convolution of length M >= 2N-1 (a power of two), done with the power-of-two
// kernels: a = pad(x.*w) -> FFT_M -> .* B -> inverse FFT_M -> .* w / M, where B = FFT_M(chirp) is
// part of the plan.  The inverse transform is conj -> forward -> conj.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bluestein_pre_kernel(const float2* __restrict__ in, float2* __restrict__ a, const float2* __restrict__ w,
                                                            unsigned n, unsigned log2m, size_t frames, int conj_in) {
  const size_t total = frames << log2m;
  const unsigned mask = (1u << log2m) - 1u;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256) {
    const size_t f = idx >> log2m;
    const unsigned j = (unsigned)idx & mask;
    float2 v = make_float2(0.0f, 0.0f);
    if (j < n) {
      v = ld_stream(in + f * n + j);
      if (conj_in) v.y = -v.y;
      v = cx_mul(v, __ldg(w + j));
    }
    a[idx] = v;
  }
}
__global__ void __launch_bounds__(256) bluestein_mul_kernel(float2* __restrict__ a, const float2* __restrict__ bspec, unsigned log2m, size_t total) {
  const unsigned mask = (1u << log2m) - 1u;
  for (size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (size_t)gridDim.x * 256)
    a[idx] = cx_mul(a[idx], __ldg(bspec + ((unsigned)idx & mask)));
}
__global__ void __launch_bounds__(256) bluestein_post_kernel(const float2* __restrict__ c, float2* __restrict__ out, const float2* __restrict__ w,
                                                             unsigned n, unsigned log2m, float inv_m, int conj_out, int do_scale, float scale) {
  const size_t f = blockIdx.y;
  const unsigned k = blockIdx.x * 256 + threadIdx.x;
  if (k >= n) return;
  float2 y = cx_mul(c[(f << log2m) + k], __ldg(w + k));
  y.x *= inv_m; y.y *= inv_m;
  if (conj_out) y.y = -y.y;
  if (do_scale) y = cx_scale_exact(y, scale);
  st_stream(out + f * n + k, y);
}

void launch_bluestein_pre(const float2* in, float2* a, const float2* w, size_t n, unsigned log2m, size_t frames, bool conj_in, int sm_count,
                          cudaStream_t st) {
  const size_t total = frames << log2m;
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)sm_count * 16;
  bluestein_pre_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(in, a, w, (unsigned)n, log2m, frames, conj_in);
}
void launch_bluestein_mul(float2* a, const float2* bspec, unsigned log2m, size_t frames, int sm_count, cudaStream_t st) {
  const size_t total = frames << log2m;
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)sm_count * 16;
  bluestein_mul_kernel<<<(unsigned)(g < cap ? g : cap), 256, 0, st>>>(a, bspec, log2m, total);
}
void launch_bluestein_post(const float2* c, float2* out, const float2* w, size_t n, unsigned log2m, size_t frames, bool conj_out,
                           bool do_scale, float scale, cudaStream_t st) {
  constexpr size_t kMaxY = 32768;
  const float inv_m = 1.0f / (float)(1ull << log2m);
  for (size_t f0 = 0; f0 < frames; f0 += kMaxY) {
    const size_t fc = frames - f0 < kMaxY ? frames - f0 : kMaxY;
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)fc);
    bluestein_post_kernel<<<grid, 256, 0, st>>>(c + (f0 << log2m), out + f0 * n, w, (unsigned)n, log2m, inv_m, conj_out, do_scale, scale);
  }
}

}  // namespace ae
