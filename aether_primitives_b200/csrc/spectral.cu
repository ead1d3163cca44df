// spectral.cu — SURVEY §8(f) "next" rows, built on the K2 device FFT (sm_100a):
//   spectrogram : the compute core of util::plot::waterfall / spectrum (src/util/plot.rs:46-68,
//                 :109-130): per fft_len chunk  vec_rfft(Scale::SN) -> vec_mirror -> |.| -> DB::from
//                 (src/util/mod.rs:26-34), fused into the FFT kernel's epilogue: 8 B in, 4 B out.
//   correlate   : the frequency-domain correlator the crate benchmarks (benches/benches.rs:410-416):
//                 vec_rfft(s) -> vec_mul(sig) -> vec_rifft(s) per frame in ONE kernel; the spectrum
//                 never leaves registers (forward output layout == inverse input layout).
#include "chain_x2.cuh"
#include "fft_device.cuh"
#include "internal.h"

namespace ae {

// CTA shape: TH threads (more when one frame needs them).  The spectrogram runs ONE transform per frame
// and is latency-bound at 16 warps/SM: 256-thread CTAs at <= 64 registers (32 warps/SM) measured
// 382 vs 318 Gsamples/s.  The correlator keeps two transforms' worth of state and is faster with
// 128-thread CTAs at <= 128 registers (262 vs 243 Gsamples/s).
template <int N, int TH>
struct SpecLaunch {
  static constexpr int T = FftCfg<N>::T;
  static constexpr int F = T >= TH ? 1 : (TH / T);
  static constexpr int THREADS = F * T;
  static constexpr int MINB = THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1);
  static constexpr size_t SMEM = (size_t)F * FftCfg<N>::SMEM_ELEMS * sizeof(float2);
};

// level of one (already scaled) bin: c.norm() = hypot (num-complex), then 10*log10 when use_db
// Fast path when |y|^2 neither overflows nor loses bits to underflow: sqrt(p) is within 2 ulp of hypot
// and 10 log10|y| = 5 log10(p) = 1.50515 log2(p) on the SFU (|error| < 2e-6 dB); otherwise the
// library hypotf/log10f evaluate it (incl. |y| = 0 -> -inf, as DB::from(0.0) gives).
__device__ __forceinline__ float level_of(float2 y, int use_db) {
  const float p = fmaf(y.x, y.x, y.y * y.y);
  if (p > 1e-30f && p < 1e30f) return use_db ? 1.5051499783199058f * sfu_lg2(p) : sfu_sqrt(p);   // MUFU, <= 2 ulp
  const float nrm = hypotf(y.x, y.y);
  return use_db ? 10.0f * log10f(nrm) : nrm;
}

template <int N, bool INV>
__global__ void __launch_bounds__(SpecLaunch<N, 256>::THREADS, SpecLaunch<N, 256>::MINB)
spectrogram_kernel(const float2* __restrict__ in, size_t n_samples, float* __restrict__ levels, const float2* __restrict__ tw,
                   size_t frames, float scale, int use_db) {
  using C = FftCfg<N>;
  using LC = SpecLaunch<N, 256>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  const int f = threadIdx.x / C::T;
  const int t = threadIdx.x % C::T;
  for (size_t frame = (size_t)blockIdx.x * LC::F + f; frame < frames; frame += (size_t)gridDim.x * LC::F) {
    const size_t base = frame * N;
    float2 x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const size_t g = base + t + m * C::T;
      x[m] = g < n_samples ? ld_stream(in + g) : make_float2(0.0f, 0.0f);  // zero padding of the last chunk (:52-58)
    }
    fft_frame<N, INV>(x, smem + f * C::SMEM_ELEMS, tw, t, f);
    float* dst = levels + base;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int pos = t + m * C::T;
      const float2 y = cx_scale_exact(x[m], scale);       // Scale::SN of vec_rfft
      __stcs(dst + ((pos + N / 2) & (N - 1)), level_of(y, use_db));  // vec_mirror: swap halves
    }
    if (C::NP > 1) frame_sync<C::T>(f);
  }
}

// fallback epilogue for lengths without a power-of-two kernel: mirror + level of an already
// transformed and scaled buffer (odd len: the last element of each chunk stays, src/vecops.rs:157-161)
__global__ void __launch_bounds__(256) levels_kernel(const float2* __restrict__ spec, float* __restrict__ levels, size_t total, size_t n,
                                                     int use_db) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t fr = i / n, k = i % n, mid = n / 2;
  size_t src = k;
  if (k < mid) src = k + mid;
  else if (k < 2 * mid) src = k - mid;
  levels[i] = level_of(spec[fr * n + src], use_db);
}

template <int N>
static void launch_spec_n(const float2* in, size_t n_samples, float* levels, const float2* tw, size_t frames, bool inverse, float scale,
                          int use_db, cudaStream_t st) {
  using LC = SpecLaunch<N, 256>;
  const size_t want = (frames + LC::F - 1) / LC::F;
  auto launch = [&](auto kern) {
    const size_t resident = resident_ctas((const void*)kern, LC::THREADS, LC::SMEM);
    kern<<<(unsigned)(want < resident ? want : resident), LC::THREADS, LC::SMEM, st>>>(in, n_samples, levels, tw, frames, scale, use_db);
  };
  if (inverse) launch(spectrogram_kernel<N, true>);
  else launch(spectrogram_kernel<N, false>);
}

void launch_spectrogram(const float2* in, size_t n_samples, float* levels, size_t n, size_t frames, const float2* tw, bool inverse,
                        float scale, int use_db, cudaStream_t st) {
  if (frames == 0) return;
  switch (n) {
#define AE_CASE(NN) case NN: launch_spec_n<NN>(in, n_samples, levels, tw, frames, inverse, scale, use_db, st); break;
    AE_CASE(16) AE_CASE(32) AE_CASE(64) AE_CASE(128) AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096) AE_CASE(8192)
#undef AE_CASE
    default: note_unsupported_launch("spectral kernels: FFT length must be a power of two in 64..4096");
  }
}
bool spectral_supported(size_t n) { return n >= 16 && n <= 8192 && (n & (n - 1)) == 0; }

void launch_levels(const float2* spec, float* levels, size_t total, size_t n, int use_db, cudaStream_t st) {
  if (total) levels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(spec, levels, total, n, use_db);
}

// -------------------------------------------------------------------------------------------------
// correlator: per frame  x <- bwd( fwd(x)*s1 .* sig )*s2   (benches/benches.rs:410-416)
// FWD_INV: exponent sign of Cfft::fwd is + (compat=reference); the backward transform uses the other.
// -------------------------------------------------------------------------------------------------
template <int N, bool FWD_INV>
__global__ void __launch_bounds__(SpecLaunch<N, 128>::THREADS, SpecLaunch<N, 128>::MINB)
correlate_kernel(float2* __restrict__ data, const float2* __restrict__ sig, const float2* __restrict__ tw, size_t frames, float scale,
                 int do_scale) {
  using C = FftCfg<N>;
  using LC = SpecLaunch<N, 128>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* sm = reinterpret_cast<float2*>(smem_raw) + (threadIdx.x / C::T) * C::SMEM_ELEMS;
  const int f = threadIdx.x / C::T;
  const int t = threadIdx.x % C::T;
  for (size_t frame = (size_t)blockIdx.x * LC::F + f; frame < frames; frame += (size_t)gridDim.x * LC::F) {
    float2* p = data + frame * N;
    float2 x[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) x[m] = ld_stream(p + t + m * C::T);
    fft_frame<N, FWD_INV>(x, sm, tw, t, f);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      if (do_scale) x[m] = cx_scale_exact(x[m], scale);                 // Scale of vec_rfft
      x[m] = cx_mul_exact(x[m], __ldg(sig + t + m * C::T));             // vec_mul: unfused arithmetic
    }
    if (C::NP > 1) frame_sync<C::T>(f);
    fft_frame<N, !FWD_INV>(x, sm, tw, t, f);
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      if (do_scale) x[m] = cx_scale_exact(x[m], scale);                 // Scale of vec_rifft
      st_stream(p + t + m * C::T, x[m]);
    }
    if (C::NP > 1) frame_sync<C::T>(f);
  }
}

// Same chain for 1024-point frames on the packed one-exchange transform of chain_x2.cuh: a warp per frame, the forward
// result (thread t: bins t + 32 c) is the inverse transform's input layout, so scale, multiply and the second transform
// stay in registers; one shared-memory exchange per transform instead of two, no barrier between warps.
template <bool FWD_INV, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
correlate_x2_kernel(float2* __restrict__ data, const float2* __restrict__ sig, const float2* __restrict__ twtab, size_t frames, float scale,
                    int do_scale) {
  constexpr int N = 1024;
  using XC = X2Cfg<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* sg = reinterpret_cast<float2*>(smem_raw);                               // N cf32: the reference signal's spectrum operand
  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  float2* ex = sg + N + (size_t)warp * (2 * XC::EX_ELEMS);
  for (int i = threadIdx.x; i < N; i += 32 * WARPS) sg[i] = __ldg(sig + i);
  X2Tw tw;
  x2_load_twiddles<N>(tw, twtab, t);
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * WARPS;
  for (size_t frame = (size_t)blockIdx.x * WARPS + warp; frame < frames; frame += stride) {
    float2* p = data + frame * N + t;
    if (frame + stride < frames) {   // next frame of this warp towards L2 (plain loads have no look-ahead)
      const char* pf = reinterpret_cast<const char*>(data + (frame + stride) * N) + 128 * t;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 4096));
    }
    float2 y[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = ld_stream(p + 32 * j);
    cx2 v[16];
    x2_stage1_regs<FWD_INV>(v, y, tw);
    c2dft16<FWD_INV>(v);
    x2_second_stage<N, FWD_INV, false, false>(v, y, ex, tw, t);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (do_scale) y[c] = cx_scale_exact(y[c], scale);                 // Scale of vec_rfft
      y[c] = cx_mul_exact(y[c], sg[t + 32 * c]);                        // vec_mul: unfused arithmetic
    }
    x2_syncwarp();
    x2_stage1_regs<!FWD_INV>(v, y, tw);
    c2dft16<!FWD_INV>(v);
    x2_second_stage<N, !FWD_INV, false, false>(v, y, ex, tw, t);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (do_scale) y[c] = cx_scale_exact(y[c], scale);                 // Scale of vec_rifft
      st_stream(p + 32 * c, y[c]);
    }
    x2_syncwarp();
  }
}
void launch_correlate_x2(float2* data, const float2* sig, size_t frames, const float2* x2tw, bool fwd_inverse, float scale, int do_scale,
                         cudaStream_t st) {
  if (frames == 0) return;
  constexpr int WARPS = 12;
  using XC = X2Cfg<1024>;
  const size_t smem = (size_t)1024 * sizeof(float2) + (size_t)WARPS * 2 * XC::EX_ELEMS * sizeof(float2);
  const size_t want = (frames + WARPS - 1) / WARPS;
  auto launch = [&](auto kern) {
    const size_t resident = resident_ctas((const void*)kern, 32 * WARPS, smem);
    kern<<<(unsigned)(want < resident ? want : resident), 32 * WARPS, smem, st>>>(data, sig, x2tw, frames, scale, do_scale);
  };
  if (fwd_inverse) launch(correlate_x2_kernel<true, WARPS>);
  else launch(correlate_x2_kernel<false, WARPS>);
}

template <int N>
static void launch_corr_n(float2* data, const float2* sig, const float2* tw, size_t frames, bool fwd_inverse, float scale, int do_scale,
                          cudaStream_t st) {
  using LC = SpecLaunch<N, 128>;
  const size_t want = (frames + LC::F - 1) / LC::F;
  auto launch = [&](auto kern) {
    const size_t resident = resident_ctas((const void*)kern, LC::THREADS, LC::SMEM);
    kern<<<(unsigned)(want < resident ? want : resident), LC::THREADS, LC::SMEM, st>>>(data, sig, tw, frames, scale, do_scale);
  };
  if (fwd_inverse) launch(correlate_kernel<N, true>);
  else launch(correlate_kernel<N, false>);
}

void launch_correlate(float2* data, const float2* sig, size_t n, size_t frames, const float2* tw, bool fwd_inverse, float scale,
                      int do_scale, cudaStream_t st) {
  if (frames == 0) return;
  switch (n) {
#define AE_CASE(NN) case NN: launch_corr_n<NN>(data, sig, tw, frames, fwd_inverse, scale, do_scale, st); break;
    AE_CASE(16) AE_CASE(32) AE_CASE(64) AE_CASE(128) AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096) AE_CASE(8192)
#undef AE_CASE
    default: note_unsupported_launch("spectral kernels: FFT length must be a power of two in 64..4096");
  }
}

}  // namespace ae
