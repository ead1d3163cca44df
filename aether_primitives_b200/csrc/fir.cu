// fir.cu — K3 direct-form FIR and K4 overlap-save FIR (sm_100a).
//
// The reference's src/fir.rs:1-22 is a constructor-only stub; the filter is the
// textbook definition  y[n] = sum_{k<T} h[k] x[n-k],  output length = input length, with either
// a carried history of T-1 samples (streaming) or zero state at the start of every frame.
// Parity tier T1: EVM against the f64 direct form (tests/test_gpu_fir.py).
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "chain_x2.cuh"
#include "fft_device.cuh"
#include "internal.h"

namespace ae {

// -------------------------------------------------------------------------------------------------
// K3 direct form.  FP32-bound: 4 FFMA per tap per output (8*T flop/sample) against 16 B/sample.
// Each thread owns S = 8 consecutive outputs and slides an 8-register window over the input
// (one new cf32 from shared memory per tap), so the inner loop is 32 FFMA : 2 LDS.
// The tap loop is unrolled by 8 so the window rotation is pure register renaming.
// -------------------------------------------------------------------------------------------------
constexpr int kFirS = 8;
constexpr int kFirThreads = 256;
constexpr int kFirTile = kFirS * kFirThreads;  // 2048 outputs per CTA

__device__ __forceinline__ int fir_pad(int i) { return i + (i >> 3); }

__global__ void __launch_bounds__(kFirThreads)
fir_direct_kernel(const float2* __restrict__ x, float2* __restrict__ y, size_t n, const float2* __restrict__ taps, int tp,
                  const float2* __restrict__ history, size_t frame_len) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* hs = reinterpret_cast<float2*>(smem_raw);  // tp taps
  float2* xs = hs + tp;                              // padded tile: kFirTile + tp inputs
  const long long tile0 = (long long)blockIdx.x * kFirTile;
  const int n_in = kFirTile + tp;
  for (int i = threadIdx.x; i < tp; i += kFirThreads) hs[i] = __ldg(taps + i);
  // tile element i <-> global sample g = tile0 - tp + i
  for (int i = threadIdx.x; i < n_in; i += kFirThreads) {
    const long long g = tile0 - tp + i;
    float2 v = make_float2(0.0f, 0.0f);
    if (g >= 0) { if ((size_t)g < n) v = ld_stream(x + g); }
    else if (history && g >= -(long long)(tp - 1)) v = __ldg(history + (tp - 1) + g);
    xs[fir_pad(i)] = v;
  }
  __syncthreads();
  const long long n0 = tile0 + (long long)threadIdx.x * kFirS;
  if ((size_t)n0 >= n) return;
  // kz = samples of this thread's frame that precede n0 (frame_len % 8 == 0 guaranteed by the host):
  // tap k of output n0+s reads x[n0+s-k], which lies before the frame start iff k - s > kz
  int kz = 0x7fffffff;
  if (frame_len) {
    const long long in_frame = n0 % (long long)frame_len;
    kz = in_frame < (long long)tp ? (int)in_frame : 0x7fffffff;
  }
  const int base = tp + threadIdx.x * kFirS;  // tile index of x[n0]
  float2 acc[kFirS], w[kFirS];
#pragma unroll
  for (int s = 0; s < kFirS; ++s) {
    acc[s] = make_float2(0.0f, 0.0f);
    w[s] = xs[fir_pad(base + s)];  // x[n0+s] belongs to the frame by construction
  }
  for (int kb = 0; kb < tp; kb += kFirS) {
#pragma unroll
    for (int kk = 0; kk < kFirS; ++kk) {
      const float2 h = hs[kb + kk];
#pragma unroll
      for (int s = 0; s < kFirS; ++s) cx_fma(acc[s], h, w[(s - kk) & (kFirS - 1)]);
      // slide: x[n0 - k - 1] replaces the oldest window entry
      const int k1 = kb + kk + 1;
      float2 v = xs[fir_pad(base - k1)];
      if (k1 > kz) v = make_float2(0.0f, 0.0f);  // x[n0 - k1] belongs to the previous frame
      w[(kFirS - 1 - kk) & (kFirS - 1)] = v;
    }
  }
  if ((size_t)(n0 + kFirS) <= n && ((uintptr_t)(y + n0) % 16) == 0) {
    float4* o = reinterpret_cast<float4*>(y + n0);
#pragma unroll
    for (int s = 0; s < kFirS; s += 2) st_stream(o + s / 2, make_float4(acc[s].x, acc[s].y, acc[s + 1].x, acc[s + 1].y));
  } else {
#pragma unroll
    for (int s = 0; s < kFirS; ++s)
      if ((size_t)(n0 + s) < n) y[n0 + s] = acc[s];
  }
}

// -------------------------------------------------------------------------------------------------
// K3, packed variant.  sm_100a has 2-wide FP32 FMAs (PTX fma.rn.f32x2, SASS FFMA2): the same FLOP
// rate as FFMA for HALF the issue slots.  A complex MAC  acc += h*x  is two of them,
//     acc = fma2( (xr, xi), (hr, hr), acc );   acc = fma2( (xi, xr), (-hi, hi), acc );
// with the tap pairs (hr,hr,-hi,hi) prepared once in shared memory (one 16-byte broadcast read per tap)
// and the swapped sample (xi, xr) built once per loaded sample and reused by its 8 outputs.  The loop
// is then bound by the FMA pipe itself instead of by instruction issue.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  return ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// CTAPS: the tap pairs come from the kernel parameters (constant bank -> uniform registers: `FFMA2 R, R, UR, R`), so a
// packed FMA reads two vector-register pairs instead of three.  Measured at 64 taps: the three-register form is bound by
// register-file bandwidth at 74 Gsamples/s (57 % of what the FMA pipe allows), whatever the rest of the loop does.
template <int NT>
struct FirTapsC { ulonglong2 h[NT]; };                        // (hr, hr), (-hi, hi) per tap: 2 KB (128 taps) or 16 KB (1024) of parameters
constexpr int kFirCTapsSmall = 128, kFirCTapsLarge = 1024;

template <bool CTAPS, int NT>
__global__ void __launch_bounds__(kFirThreads)
fir_direct_x2_kernel(const float2* __restrict__ x, float2* __restrict__ y, size_t n, const float2* __restrict__ taps, int tp,
                     const float2* __restrict__ history, size_t frame_len, const __grid_constant__ FirTapsC<NT> tc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* hs = reinterpret_cast<float4*>(smem_raw);          // tp entries (hr, hr, -hi, hi); unused with CTAPS
  float2* xs = reinterpret_cast<float2*>(hs + (CTAPS ? 0 : tp));   // padded tile: kFirTile + tp inputs
  const long long tile0 = (long long)blockIdx.x * kFirTile;
  const int n_in = kFirTile + tp;
  if (!CTAPS) {
    for (int i = threadIdx.x; i < tp; i += kFirThreads) {
      const float2 h = __ldg(taps + i);
      hs[i] = make_float4(h.x, h.x, -h.y, h.y);
    }
  }
  for (int i = threadIdx.x; i < n_in; i += kFirThreads) {
    const long long g = tile0 - tp + i;
    float2 v = make_float2(0.0f, 0.0f);
    if (g >= 0) { if ((size_t)g < n) v = ld_stream(x + g); }
    else if (history && g >= -(long long)(tp - 1)) v = __ldg(history + (tp - 1) + g);
    xs[fir_pad(i)] = v;
  }
  __syncthreads();
  const long long n0 = tile0 + (long long)threadIdx.x * kFirS;
  if ((size_t)n0 >= n) return;
  int kz = 0x7fffffff;
  if (frame_len) {
    const long long in_frame = n0 % (long long)frame_len;
    kz = in_frame < (long long)tp ? (int)in_frame : 0x7fffffff;
  }
  const int base = tp + threadIdx.x * kFirS;
  unsigned long long acc[kFirS], w[kFirS], wsw[kFirS];  // (re, im) accumulators; samples natural and swapped
#pragma unroll
  for (int s = 0; s < kFirS; ++s) {
    acc[s] = 0ull;
    const float2 v = xs[fir_pad(base + s)];
    w[s] = f2_pack(v.x, v.y);
    wsw[s] = f2_pack(v.y, v.x);
  }
  // tp and base are multiples of 8, so the eight samples a block of eight taps pulls in, x[n0 - kb - 1 .. n0 - kb - 8], sit in
  // ONE padded group of the tile: one pointer per block, compile-time offsets inside it.  Threads whose window never
  // reaches back before their frame (all of them when streaming) skip the per-tap zeroing test.
  const ulonglong2* hp = reinterpret_cast<const ulonglong2*>(hs);      // (hr, hr), (-hi, hi): one 128-bit broadcast read per tap
  auto taps_loop = [&](auto check_tag) {
    constexpr bool CHECK = decltype(check_tag)::value;
    for (int kb = 0; kb < tp; kb += kFirS) {
      const float2* xp = xs + fir_pad(base - kb - kFirS);
#pragma unroll
      for (int kk = 0; kk < kFirS; ++kk) {
        const ulonglong2 h = CTAPS ? tc.h[kb + kk] : hp[kb + kk];
#pragma unroll
        for (int s = 0; s < kFirS; ++s) {
          const int j = (s - kk) & (kFirS - 1);
          acc[s] = f2_fma(w[j], h.x, acc[s]);
          acc[s] = f2_fma(wsw[j], h.y, acc[s]);
        }
        float2 v = xp[kFirS - 1 - kk];                     // x[n0 - (kb + kk + 1)]
        if (CHECK && kb + kk + 1 > kz) v = make_float2(0.0f, 0.0f);
        const int jn = (kFirS - 1 - kk) & (kFirS - 1);
        w[jn] = f2_pack(v.x, v.y);
        wsw[jn] = f2_pack(v.y, v.x);
      }
    }
  };
  if (kz == 0x7fffffff) taps_loop(std::false_type{});
  else taps_loop(std::true_type{});
  if ((size_t)(n0 + kFirS) <= n && ((uintptr_t)(y + n0) % 16) == 0) {
    ulonglong2* o = reinterpret_cast<ulonglong2*>(y + n0);
#pragma unroll
    for (int s = 0; s < kFirS; s += 2) __stcs(o + s / 2, make_ulonglong2(acc[s], acc[s + 1]));
  } else {
#pragma unroll
    for (int s = 0; s < kFirS; ++s)
      if ((size_t)(n0 + s) < n) y[n0 + s] = make_float2(__uint_as_float((unsigned)acc[s]), __uint_as_float((unsigned)(acc[s] >> 32)));
  }
}

void launch_fir_direct(const float2* x, float2* y, size_t n, const float2* taps_padded, int tp, const float2* history,
                       size_t frame_len, int, cudaStream_t st, const float2* taps_host_padded) {
  if (n == 0) return;
  static const char* scalar = getenv("AE_FIR_SCALAR");
  static const char* no_ctaps = getenv("AE_FIR_NO_CTAPS");
  if (!scalar) {
    const size_t n_in2 = (size_t)kFirTile + tp;
    const bool ctaps = taps_host_padded && tp <= kFirCTapsLarge && !no_ctaps;
    const size_t smem2 = (ctaps ? 0 : (size_t)tp * sizeof(float4)) + (n_in2 + n_in2 / 8 + 2) * sizeof(float2);
    const unsigned grid = (unsigned)((n + kFirTile - 1) / kFirTile);
    auto fill = [&](ulonglong2* dst, int nt) {
      for (int i = 0; i < nt; ++i) {
        const float2 h = i < tp ? taps_host_padded[i] : make_float2(0.0f, 0.0f);
        const float nh = -h.y;
        unsigned hr, hi, nhi;
        memcpy(&hr, &h.x, 4); memcpy(&hi, &h.y, 4); memcpy(&nhi, &nh, 4);
        dst[i].x = ((unsigned long long)hr << 32) | hr;        // (hr, hr)
        dst[i].y = ((unsigned long long)hi << 32) | nhi;       // (-hi, hi)
      }
    };
    if (ctaps && tp <= kFirCTapsSmall) {
      FirTapsC<kFirCTapsSmall> tc;
      fill(tc.h, kFirCTapsSmall);
      fir_direct_x2_kernel<true, kFirCTapsSmall><<<grid, kFirThreads, smem2, st>>>(x, y, n, taps_padded, tp, history, frame_len, tc);
      return;
    }
    if (ctaps) {
      FirTapsC<kFirCTapsLarge> tc;
      fill(tc.h, kFirCTapsLarge);
      if (smem2 > 48 * 1024)
        cudaFuncSetAttribute(fir_direct_x2_kernel<true, kFirCTapsLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      fir_direct_x2_kernel<true, kFirCTapsLarge><<<grid, kFirThreads, smem2, st>>>(x, y, n, taps_padded, tp, history, frame_len, tc);
      return;
    }
    FirTapsC<1> none;
    memset(&none, 0, sizeof(none));
    if (smem2 > 48 * 1024) cudaFuncSetAttribute(fir_direct_x2_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    fir_direct_x2_kernel<false, 1><<<grid, kFirThreads, smem2, st>>>(x, y, n, taps_padded, tp, history, frame_len, none);
    return;
  }
  const size_t n_in = (size_t)kFirTile + tp;
  const size_t smem = ((size_t)tp + n_in + n_in / 8 + 2) * sizeof(float2);
  if (smem > 48 * 1024) cudaFuncSetAttribute(fir_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned grid = (unsigned)((n + kFirTile - 1) / kFirTile);
  fir_direct_kernel<<<grid, kFirThreads, smem, st>>>(x, y, n, taps_padded, tp, history, frame_len);
}

// -------------------------------------------------------------------------------------------------
// K4 overlap-save: one segment of NF inputs per frame slot -> FFT -> * H -> inverse FFT -> keep the
// last L = NF - T + 1 outputs.  The forward transform's register layout (position t + m*T) is
// exactly the inverse transform's input layout, so the spectrum never leaves registers.
// H already carries the 1/NF of the round trip.  Flops/sample: (2*5*NF*log2 NF + 6*NF)/L.
// -------------------------------------------------------------------------------------------------
template <int NF>
struct OsLaunch {
  static constexpr int T = FftCfg<NF>::T;
  static constexpr int F = T >= 256 ? 1 : (256 / T);
  static constexpr int THREADS = F * T;
  static constexpr size_t SMEM = (size_t)F * FftCfg<NF>::SMEM_ELEMS * sizeof(float2);
};

template <int NF>
__global__ void __launch_bounds__(OsLaunch<NF>::THREADS)
fir_os_kernel(const float2* __restrict__ x, float2* __restrict__ y, size_t n, const float2* __restrict__ H,
              const float2* __restrict__ tw, int ntaps, int hist_len, const float2* __restrict__ history, size_t frame_len,
              size_t segs_per_frame, size_t n_segs) {
  using C = FftCfg<NF>;
  using LC = OsLaunch<NF>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* sm = reinterpret_cast<float2*>(smem_raw) + (threadIdx.x / C::T) * C::SMEM_ELEMS;
  const int f = threadIdx.x / C::T;
  const int t = threadIdx.x % C::T;
  const size_t seg = (size_t)blockIdx.x * LC::F + f;
  if (seg >= n_segs) return;
  const long long L = NF - ntaps + 1;
  long long out_start, frame_start, out_end;
  if (frame_len) {
    const size_t fi = seg / segs_per_frame, si = seg % segs_per_frame;
    frame_start = (long long)(fi * frame_len);
    out_start = frame_start + (long long)si * L;
    long long fe = frame_start + (long long)frame_len;
    if (fe > (long long)n) fe = (long long)n;
    out_end = out_start + L < fe ? out_start + L : fe;
  } else {
    frame_start = 0;
    out_start = (long long)seg * L;
    out_end = out_start + L < (long long)n ? out_start + L : (long long)n;
  }
  const long long in_start = out_start - (ntaps - 1);
  // interior segment: every input exists inside the frame and every output is kept, so neither the
  // loads nor the stores need per-element bounds tests (they were 20 % of the executed instructions)
  const bool interior = in_start >= frame_start && out_end == out_start + L;
  float2 v[16];
  if (interior) {
    const float2* src = x + in_start + t;
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = ld_stream(src + m * C::T);
  } else {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const long long g = in_start + t + m * C::T;
      float2 s = make_float2(0.0f, 0.0f);
      if (g >= frame_start) { if (g < out_end) s = ld_stream(x + g); }
      else if (!frame_len && history && g >= -(long long)hist_len) s = __ldg(history + hist_len + g);
      v[m] = s;
    }
  }
  fft_frame<NF, false>(v, sm, tw, t, f);
#pragma unroll
  for (int m = 0; m < 16; ++m) v[m] = cx_mul(v[m], __ldg(H + t + m * C::T));
  frame_sync<C::T>(f);  // every thread is past its last read of sm
  fft_frame<NF, true>(v, sm, tw, t, f);
  if (interior) {
    float2* dst = y + in_start + t;
#pragma unroll
    for (int m = 0; m < 16; ++m)
      if (t + m * C::T >= ntaps - 1) st_stream(dst + m * C::T, v[m]);
  } else {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int i = t + m * C::T;
      const long long g = out_start + i - (ntaps - 1);
      if (i >= ntaps - 1 && g < out_end) st_stream(y + g, v[m]);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// K4b overlap-save for 1024-point blocks (up to 256 taps) on the transform of chain_x2.cuh: one WARP per segment,
// 1024 = 32 x 32 with one shared-memory exchange per transform (K4 needs two, and a named barrier between the two
// warps that share a segment), packed FP32 butterflies, persistent warps.  The forward transform leaves thread t with
// bins t + 32 c, which is exactly the input layout of the inverse transform, so the spectrum is multiplied by H in
// registers.  Same results as K4 up to rounding (both are checked against the f64 direct form).
// -------------------------------------------------------------------------------------------------
template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
fir_os_x2_kernel(const float2* __restrict__ x, float2* __restrict__ y, size_t n, const float2* __restrict__ H,
                 const float2* __restrict__ twtab, int ntaps, int hist_len, const float2* __restrict__ history, size_t frame_len,
                 size_t segs_per_frame, size_t n_segs) {
  constexpr int NF = 1024;
  using XC = X2Cfg<NF>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* Hs = reinterpret_cast<float2*>(smem_raw);                               // NF cf32
  const int warp = threadIdx.x >> 5, t = threadIdx.x & 31;
  float2* ex = Hs + NF + (size_t)warp * (2 * XC::EX_ELEMS);                       // two planes of EX_ELEMS float2 per warp
  for (int i = threadIdx.x; i < NF; i += 32 * WARPS) Hs[i] = __ldg(H + i);
  X2Tw tw;
  x2_load_twiddles<NF>(tw, twtab, t);
  __syncthreads();
  const long long L = NF - ntaps + 1;
  for (size_t seg = (size_t)blockIdx.x * WARPS + warp; seg < n_segs; seg += (size_t)gridDim.x * WARPS) {
    long long out_start, frame_start, out_end;
    if (frame_len) {
      const size_t fi = seg / segs_per_frame, si = seg % segs_per_frame;
      frame_start = (long long)(fi * frame_len);
      out_start = frame_start + (long long)si * L;
      long long fe = frame_start + (long long)frame_len;
      if (fe > (long long)n) fe = (long long)n;
      out_end = out_start + L < fe ? out_start + L : fe;
    } else {
      frame_start = 0;
      out_start = (long long)seg * L;
      out_end = out_start + L < (long long)n ? out_start + L : (long long)n;
    }
    const long long in_start = out_start - (ntaps - 1);
    const bool interior = in_start >= frame_start && out_end == out_start + L;
    {
      // pull this warp's NEXT segment towards L2 while the current one is transformed (plain loads have no
      // look-ahead of their own; 8 KB = 64 lines = two per lane)
      const size_t nseg = seg + (size_t)gridDim.x * WARPS;
      if (!frame_len && nseg < n_segs) {
        const long long ns = (long long)nseg * L - (ntaps - 1);
        const char* pf = reinterpret_cast<const char*>(x + (ns > 0 ? ns : 0)) + 128 * t;
        if ((long long)(ns + NF) <= (long long)n) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 4096));
        }
      }
    }
    float2 v32[32];
    if (interior) {
      const float2* src = x + in_start + t;
#pragma unroll
      for (int j = 0; j < 32; ++j) v32[j] = ld_stream(src + 32 * j);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const long long g = in_start + t + 32 * j;
        float2 s = make_float2(0.0f, 0.0f);
        if (g >= frame_start) { if (g < out_end) s = ld_stream(x + g); }
        else if (!frame_len && history && g >= -(long long)hist_len) s = __ldg(history + hist_len + g);
        v32[j] = s;
      }
    }
    cx2 v[16];
    x2_stage1_regs<false>(v, v32, tw);
    c2dft16<false>(v);
    x2_second_stage<NF, false, false, false>(v, v32, ex, tw, t);
#pragma unroll
    for (int c = 0; c < 32; ++c) v32[c] = cx_mul(v32[c], Hs[t + 32 * c]);
    x2_syncwarp();                                         // every lane is past its reads of the exchange buffer
    x2_stage1_regs<true>(v, v32, tw);
    c2dft16<true>(v);
    x2_second_stage<NF, true, false, false>(v, v32, ex, tw, t);
    if (interior) {
      float2* dst = y + in_start + t;
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (t + 32 * c >= ntaps - 1) st_stream(dst + 32 * c, v32[c]);
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int i = t + 32 * c;
        const long long g = out_start + i - (ntaps - 1);
        if (i >= ntaps - 1 && g < out_end) st_stream(y + g, v32[c]);
      }
    }
    x2_syncwarp();
  }
}

template <int WARPS>
static void launch_fir_os_x2_w(const float2* x, float2* y, size_t n, const float2* H, const float2* x2tw, size_t ntaps, const float2* history,
                               size_t frame_len, cudaStream_t st) {
  constexpr int NF = 1024;
  using XC = X2Cfg<NF>;
  const int hist_len = (int)(((ntaps + 7) / 8) * 8) - 1;
  const size_t L = NF - ntaps + 1;
  size_t segs_per_frame = 0, n_segs;
  if (frame_len) {
    segs_per_frame = (frame_len + L - 1) / L;
    n_segs = ((n + frame_len - 1) / frame_len) * segs_per_frame;
  } else {
    n_segs = (n + L - 1) / L;
  }
  const size_t smem = (size_t)NF * sizeof(float2) + (size_t)WARPS * 2 * XC::EX_ELEMS * sizeof(float2);
  const size_t resident = resident_ctas((const void*)fir_os_x2_kernel<WARPS>, 32 * WARPS, smem);
  const size_t want = (n_segs + WARPS - 1) / WARPS;
  fir_os_x2_kernel<WARPS><<<(unsigned)(want < resident ? want : resident), 32 * WARPS, smem, st>>>(
      x, y, n, H, x2tw, (int)ntaps, hist_len, history, frame_len, segs_per_frame, n_segs);
}

void launch_fir_os_x2(const float2* x, float2* y, size_t n, const float2* H, const float2* x2tw, size_t ntaps, const float2* history,
                      size_t frame_len, cudaStream_t st) {
  if (n == 0) return;
  static const char* we = getenv("AE_FIR_WARPS");
  const int warps = we ? atoi(we) : 12;   // measured at 64 taps, 2^28 samples: 8 -> 311, 10 -> 297, 12 -> 342, 16 -> 304 Gsamples/s
  if (warps >= 16) return launch_fir_os_x2_w<16>(x, y, n, H, x2tw, ntaps, history, frame_len, st);
  if (warps >= 12) return launch_fir_os_x2_w<12>(x, y, n, H, x2tw, ntaps, history, frame_len, st);
  if (warps >= 10) return launch_fir_os_x2_w<10>(x, y, n, H, x2tw, ntaps, history, frame_len, st);
  return launch_fir_os_x2_w<8>(x, y, n, H, x2tw, ntaps, history, frame_len, st);
}

bool fir_os_supported(size_t nfft) { return nfft >= 256 && nfft <= 16384 && (nfft & (nfft - 1)) == 0; }

template <int NF>
static void launch_os_n(const float2* x, float2* y, size_t n, const float2* H, const float2* tw, size_t ntaps,
                        const float2* history, int hist_len, size_t frame_len, cudaStream_t st) {
  using LC = OsLaunch<NF>;
  const size_t L = NF - ntaps + 1;
  size_t segs_per_frame = 0, n_segs;
  if (frame_len) {
    segs_per_frame = (frame_len + L - 1) / L;
    n_segs = ((n + frame_len - 1) / frame_len) * segs_per_frame;
  } else {
    n_segs = (n + L - 1) / L;
  }
  (void)resident_ctas((const void*)fir_os_kernel<NF>, LC::THREADS, LC::SMEM);   // shared-memory opt-in, once per device
  const unsigned grid = (unsigned)((n_segs + LC::F - 1) / LC::F);
  fir_os_kernel<NF><<<grid, LC::THREADS, LC::SMEM, st>>>(x, y, n, H, tw, (int)ntaps, hist_len, history, frame_len, segs_per_frame, n_segs);
}

void launch_fir_overlap_save(const float2* x, float2* y, size_t n, const float2* H, const float2* tw, size_t nfft, size_t ntaps,
                             const float2* history, size_t frame_len, cudaStream_t st) {
  if (n == 0) return;
  // history layout is shared with the direct kernel: (ntaps rounded up to 8) - 1 samples, newest last
  const int hist_len = (int)(((ntaps + 7) / 8) * 8) - 1;
  switch (nfft) {
#define AE_CASE(NN) case NN: launch_os_n<NN>(x, y, n, H, tw, ntaps, history, hist_len, frame_len, st); break;
    AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096) AE_CASE(8192) AE_CASE(16384)
#undef AE_CASE
    default: note_unsupported_launch("overlap-save FIR: block length must be a power of two in 256..16384");
  }
}

}  // namespace ae
