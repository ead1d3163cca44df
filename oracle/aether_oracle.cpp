// aether_oracle.cpp — CPU restatement of the aether_primitives cf32 hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under aether_primitives_b200/ may include,
// link, load or call this file.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker and as
// the reported CPU baseline.
//
// The reference crate (razorheadfx/aether_primitives) is Rust and cannot be built
// here (no rustc/cargo, nightly-only crate, un-vendored crates.io dependencies), so
// this is a line-by-line RESTATEMENT in C++17, compiled with
//   g++ -O3 -ffp-contract=off -fno-fast-math
// so that, like rustc, no multiply-add is ever contracted into an FMA.  Every
// function cites the reference file:line it follows.  Citations are relative to the
// reference repository root.
//
// Parity status (see DESIGN.md "Oracle pinning"):
//  * VecOps, Scale, sampling, modulation, sequence, assert_evm!: pinned against every
//    golden vector the reference's own tests hold (tests/test_oracle_golden.py).
//  * FFT: the butterflies live in rustfft ^3.0 (Cargo.toml:27), which is not vendored.
//    The reference tests pin only "unnormalised, any length, round trip = N*x" and DC
//    bins; sign convention and per-bin accuracy are PARITY UNPINNED.  This oracle is a
//    f32 mixed-radix Cooley-Tukey FFT with twiddles computed in f64 and rounded to f32
//    (rustfft's accuracy class), plus an f64 evaluation used as float truth.
//  * FIR: src/fir.rs:1-22 holds no filter at all -> PARITY UNPINNED; oracle = the
//    textbook direct form, f32 in tap order and f64 truth.
//  * AWGN: rand 0.7 ChaCha20 + rand_distr ziggurat are not vendored and the north
//    star replaces them by Philox4x32-10 + Box-Muller -> validated statistically; the
//    Philox block function is pinned by the Random123 known-answer vectors.

#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

typedef struct { float re, im; } ocf32;

enum { ORA_OK = 0, ORA_ELEN = 1, ORA_EARG = 2, ORA_EIDX = 6 };
enum { ORA_COMPAT_REFERENCE = 0, ORA_COMPAT_CORRECTED = 1 };
enum { ORA_SCALE_NONE = 0, ORA_SCALE_SN = 1, ORA_SCALE_N = 2, ORA_SCALE_X = 3 };

// ----------------------------------------------------------------------------------
// assert_evm!  (src/lib.rs:26-49)
// per element: evm = (act - ref).norm(); limit = ref.norm() * (10f64.powf(db/10) as f32)
// returns 0 = pass, 100 = "EVM limit exceeded" (bad index in *bad), ORA_ELEN, ORA_EARG
// ----------------------------------------------------------------------------------
int ora_assert_evm(const ocf32* act, size_t na, const ocf32* ref, size_t nr, double db,
                   size_t* bad) {
  if (na != nr) return ORA_ELEN;       // src/lib.rs:34
  if (!(db < 0.0)) return ORA_EARG;    // src/lib.rs:35
  const float thr = (float)std::pow(10.0, db / 10.0);
  for (size_t i = 0; i < na; ++i) {
    const float dre = act[i].re - ref[i].re, dim = act[i].im - ref[i].im;
    const float evm = hypotf(dre, dim);                       // Complex::norm == hypot
    const float limit = hypotf(ref[i].re, ref[i].im) * thr;   // src/lib.rs:38
    if (evm > limit) {
      if (bad) *bad = i;
      return 100;
    }
  }
  return ORA_OK;
}

// Power EVM the macro's DOC describes (src/lib.rs:21): 10 log10(sum|e|^2 / sum|r|^2).
double ora_evm_power_db(const ocf32* act, const ocf32* ref, size_t n) {
  double e = 0, r = 0;
  for (size_t i = 0; i < n; ++i) {
    const double dr = (double)act[i].re - ref[i].re, di = (double)act[i].im - ref[i].im;
    e += dr * dr + di * di;
    r += (double)ref[i].re * ref[i].re + (double)ref[i].im * ref[i].im;
  }
  if (e == 0) return -INFINITY;
  return 10.0 * std::log10(e / r);
}

// worst element in the macro's own (amplitude-ratio) sense, as "macro dB"
double ora_evm_macro_worst_db(const ocf32* act, const ocf32* ref, size_t n) {
  double worst = -INFINITY;
  for (size_t i = 0; i < n; ++i) {
    const float evm = hypotf(act[i].re - ref[i].re, act[i].im - ref[i].im);
    const float rn = hypotf(ref[i].re, ref[i].im);
    if (evm == 0) continue;
    const double d = 10.0 * std::log10((double)evm / (double)rn);
    if (d > worst) worst = d;
  }
  return worst;
}

// ----------------------------------------------------------------------------------
// VecOps  (src/vecops.rs:94-182; num-complex 0.2 operator formulas)
// ----------------------------------------------------------------------------------
void ora_vec_scale(ocf32* v, size_t n, float s) {  // :94-97  Complex::scale
  for (size_t i = 0; i < n; ++i) { v[i].re = v[i].re * s; v[i].im = v[i].im * s; }
}
int ora_vec_mul(ocf32* v, size_t n, const ocf32* o, size_t no) {  // :99-112
  if (n != no) return ORA_ELEN;
  for (size_t i = 0; i < n; ++i) {
    const float ar = v[i].re, ai = v[i].im, br = o[i].re, bi = o[i].im;
    const float re = ar * br - ai * bi;
    const float im = ar * bi + ai * br;
    v[i].re = re; v[i].im = im;
  }
  return ORA_OK;
}
int ora_vec_div(ocf32* v, size_t n, const ocf32* o, size_t no) {  // :114-125
  if (n != no) return ORA_ELEN;
  for (size_t i = 0; i < n; ++i) {
    const float ar = v[i].re, ai = v[i].im, br = o[i].re, bi = o[i].im;
    const float nrm = br * br + bi * bi;      // norm_sqr
    const float re = ar * br + ai * bi;
    const float im = ai * br - ar * bi;
    v[i].re = re / nrm; v[i].im = im / nrm;
  }
  return ORA_OK;
}
void ora_vec_conj(ocf32* v, size_t n) {  // :127-130
  for (size_t i = 0; i < n; ++i) v[i].im = -v[i].im;
}
int ora_vec_add(ocf32* v, size_t n, const ocf32* o, size_t no) {  // :132-142
  if (n != no) return ORA_ELEN;
  for (size_t i = 0; i < n; ++i) { v[i].re = v[i].re + o[i].re; v[i].im = v[i].im + o[i].im; }
  return ORA_OK;
}
int ora_vec_sub(ocf32* v, size_t n, const ocf32* o, size_t no) {  // :144-155
  if (n != no) return ORA_ELEN;
  for (size_t i = 0; i < n; ++i) { v[i].re = v[i].re - o[i].re; v[i].im = v[i].im - o[i].im; }
  return ORA_OK;
}
void ora_vec_mirror(ocf32* v, size_t n) {  // :157-161  (odd n: last element untouched)
  const size_t mid = n / 2;
  for (size_t x = 0; x < mid; ++x) { ocf32 t = v[x]; v[x] = v[x + mid]; v[x + mid] = t; }
}
int ora_vec_clone(ocf32* v, size_t n, const ocf32* o, size_t no) {  // :163-172
  if (n != no) return ORA_ELEN;
  std::memcpy(v, o, n * sizeof(ocf32));
  return ORA_OK;
}
void ora_vec_zero(ocf32* v, size_t n) {  // :174-177
  for (size_t i = 0; i < n; ++i) { v[i].re = 0.0f; v[i].im = 0.0f; }
}

// ----------------------------------------------------------------------------------
// Scale  (src/fft.rs:6-37)
// ----------------------------------------------------------------------------------
float ora_scale_factor(int kind, size_t n, float x) {
  switch (kind) {
    case ORA_SCALE_SN: return 1.0f / sqrtf((float)n);  // (len as f32).sqrt().recip()
    case ORA_SCALE_N: return 1.0f / (float)n;          // (len as f32).recip()
    case ORA_SCALE_X: return x;
    default: return 1.0f;
  }
}
void ora_scale(int kind, float x, ocf32* v, size_t n) {
  if (kind == ORA_SCALE_NONE) return;  // :24
  ora_vec_scale(v, n, ora_scale_factor(kind, n, x));
}

}  // extern "C"

// ----------------------------------------------------------------------------------
// FFT  (src/fft.rs:147-230 calls rustfft ^3.0: unnormalised, any length)
// Mixed-radix decimation-in-time, radix 4 / 2 / 3 / 5 / generic prime, twiddles from
// f64 cos/sin rounded to T.
// ----------------------------------------------------------------------------------
namespace {

template <typename T>
struct FftPlan {
  size_t n;
  std::vector<std::complex<T>> tw;  // tw[k] = exp(-2 pi i k / n)
  std::vector<size_t> radices;
  explicit FftPlan(size_t n_) : n(n_), tw(n_ ? n_ : 1) {
    for (size_t k = 0; k < n; ++k) {
      const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
      tw[k] = std::complex<T>((T)std::cos((double)a), (T)std::sin((double)a));
    }
    size_t m = n;
    while (m > 1) {
      size_t p;
      if (m % 4 == 0) p = 4;
      else if (m % 2 == 0) p = 2;
      else if (m % 3 == 0) p = 3;
      else if (m % 5 == 0) p = 5;
      else {
        p = 7;
        while (m % p) { p += 2; if (p * p > m) { p = m; break; } }
      }
      radices.push_back(p);
      m /= p;
    }
  }
};

template <typename T>
inline std::complex<T> cmul(const std::complex<T>& a, const std::complex<T>& b) {
  return std::complex<T>(a.real() * b.real() - a.imag() * b.imag(),
                         a.real() * b.imag() + a.imag() * b.real());
}

// out[0..n) = DFT of in[0], in[stride], ...; sign<0: exp(-), sign>0: exp(+)
template <typename T>
void fft_rec(const FftPlan<T>& pl, const std::complex<T>* in, std::complex<T>* out, size_t n,
             size_t stride, size_t level, int sign, std::complex<T>* scratch) {
  if (n == 1) { out[0] = in[0]; return; }
  const size_t p = pl.radices[level];
  const size_t m = n / p;
  for (size_t r = 0; r < p; ++r)
    fft_rec(pl, in + r * stride, out + r * m, m, stride * p, level + 1, sign, scratch);
  const size_t tws = pl.n / n;  // twiddle stride for this level: W_n^k = tw[k*tws]
  auto W = [&](size_t idx) {
    std::complex<T> w = pl.tw[idx % pl.n];
    return sign < 0 ? w : std::conj(w);
  };
  if (p == 2) {
    for (size_t k = 0; k < m; ++k) {
      const std::complex<T> a = out[k], b = cmul(out[k + m], W(k * tws));
      out[k] = a + b;
      out[k + m] = a - b;
    }
  } else if (p == 4) {
    for (size_t k = 0; k < m; ++k) {
      const std::complex<T> a = out[k];
      const std::complex<T> b = cmul(out[k + m], W(k * tws));
      const std::complex<T> c = cmul(out[k + 2 * m], W(2 * k * tws));
      const std::complex<T> d = cmul(out[k + 3 * m], W(3 * k * tws));
      const std::complex<T> s0 = a + c, s1 = a - c, s2 = b + d, s3 = b - d;
      // -i*s3 for forward (sign<0), +i*s3 for backward
      const std::complex<T> js3 = sign < 0 ? std::complex<T>(s3.imag(), -s3.real())
                                           : std::complex<T>(-s3.imag(), s3.real());
      out[k] = s0 + s2;
      out[k + m] = s1 + js3;
      out[k + 2 * m] = s0 - s2;
      out[k + 3 * m] = s1 - js3;
    }
  } else if (p == 3 || p == 5 || p == 7 || p == 11 || p == 13) {
    // rustfft's small odd butterflies (Butterfly3 / Butterfly5 / ... of rustfft ^3.0, restated from its published
    // algorithm; the crate is not vendored): with a_r = x_r + x_{p-r}, b_r = x_r - x_{p-r},
    //   X[q], X[p-q] = (x0 + sum_r Re(w^(rq)) a_r)  +/-  i-rotated (sum_r Im(w^(rq)) b_r),
    // every product and sum rounded separately, left to right (this file is built with -ffp-contract=off).
    // For a constant input the cosine sum cancels to exactly 0 in f32 (cos(4 pi/5) == -(1/2 + cos(2 pi/5)) holds for
    // the rounded constants too), which is what lets the reference's N = 100 round-trip tests
    // (src/vecops.rs:445-463) pass assert_evm! at -80 = equality to the bit.
    std::complex<T>* t = scratch;  // p entries
    const size_t wp = pl.n / p;    // W_p^q = tw[q*wp]
    const size_t hh = (p - 1) / 2;
    std::complex<T> a[7], b[7], o[13];
    for (size_t k = 0; k < m; ++k) {
      for (size_t r = 0; r < p; ++r) t[r] = cmul(out[k + r * m], W(r * k * tws));
      std::complex<T> sum = t[0];
      for (size_t r = 1; r <= hh; ++r) { a[r] = t[r] + t[p - r]; b[r] = t[r] - t[p - r]; sum += a[r]; }
      o[0] = sum;
      for (size_t q = 1; q <= hh; ++q) {
        T ere = t[0].real(), eim = t[0].imag(), fre = 0, fim = 0;
        for (size_t r = 1; r <= hh; ++r) {
          const std::complex<T> w = W(((r * q) % p) * wp);
          ere = ere + w.real() * a[r].real();
          eim = eim + w.real() * a[r].imag();
          fre = fre + w.imag() * b[r].imag();
          fim = fim + w.imag() * b[r].real();
        }
        o[q] = std::complex<T>(ere - fre, eim + fim);
        o[p - q] = std::complex<T>(ere + fre, eim - fim);
      }
      for (size_t q = 0; q < p; ++q) out[k + q * m] = o[q];
    }
  } else {
    std::complex<T>* t = scratch;  // p entries
    const size_t wp = pl.n / p;    // W_p^q = tw[q*wp]
    for (size_t k = 0; k < m; ++k) {
      for (size_t r = 0; r < p; ++r) t[r] = cmul(out[k + r * m], W(r * k * tws));
      for (size_t q = 0; q < p; ++q) {
        std::complex<T> acc = t[0];
        for (size_t r = 1; r < p; ++r) acc += cmul(t[r], W(((r * q) % p) * wp));
        out[k + q * m] = acc;
      }
    }
  }
}

template <typename T>
void fft_exec(const FftPlan<T>& pl, const std::complex<T>* in, std::complex<T>* out, int sign) {
  std::vector<std::complex<T>> scratch(pl.n + 8);
  fft_rec(pl, in, out, pl.n, 1, 0, sign, scratch.data());
}

struct PlanCache {
  std::vector<FftPlan<float>*> f;
  std::vector<FftPlan<double>*> d;
  FftPlan<float>* getf(size_t n) {
    for (auto* p : f) if (p->n == n) return p;
    f.push_back(new FftPlan<float>(n));
    return f.back();
  }
  FftPlan<double>* getd(size_t n) {
    for (auto* p : d) if (p->n == n) return p;
    d.push_back(new FftPlan<double>(n));
    return d.back();
  }
};
thread_local PlanCache g_plans;

// exponent sign of Cfft::{fwd,bwd}.  src/fft.rs:148 plans "fwd" with
// FFTplanner::new(true) and rustfft 3.0's ctor argument is `inverse`, so the reference's
// forward transform uses exp(+2 pi i nk/N).  compat=corrected swaps it back.
inline int dir_sign(int dir_bwd, int compat) {
  const int fwd_sign = (compat == ORA_COMPAT_REFERENCE) ? +1 : -1;
  return dir_bwd ? -fwd_sign : fwd_sign;
}

}  // namespace

extern "C" {

// raw transforms: sign = -1 -> exp(-2 pi i nk/N), +1 -> exp(+...).  in != out.
int ora_fft_raw(const ocf32* in, ocf32* out, size_t n, int sign) {
  if (n == 0) return ORA_EARG;
  FftPlan<float>* pl = g_plans.getf(n);
  fft_exec<float>(*pl, reinterpret_cast<const std::complex<float>*>(in),
                  reinterpret_cast<std::complex<float>*>(out), sign);
  return ORA_OK;
}
// f64 truth of the same transform on the f32 input; out = n complex doubles
int ora_fft_raw_f64(const ocf32* in, double* out, size_t n, int sign) {
  if (n == 0) return ORA_EARG;
  FftPlan<double>* pl = g_plans.getd(n);
  std::vector<std::complex<double>> tmp(n);
  for (size_t i = 0; i < n; ++i) tmp[i] = std::complex<double>(in[i].re, in[i].im);
  fft_exec<double>(*pl, tmp.data(), reinterpret_cast<std::complex<double>*>(out), sign);
  return ORA_OK;
}

// Cfft::{fwd,bwd,ifwd,ibwd,tfwd,tbwd}  (src/fft.rs:162-230): copy input to tmp, transform
// tmp -> out, then Scale::scale(out) as a SEPARATE rounding step.  `howmany` frames of
// `n` (batching is new surface; per frame this is exactly the reference sequence).
// in may equal out (ifwd/ibwd).
int ora_cfft_exec(const ocf32* in, size_t in_len, ocf32* out, size_t n, size_t howmany,
                  int dir_bwd, int scale_kind, float scale_x, int compat) {
  if (n == 0 || in_len != n * howmany) return ORA_ELEN;  // :163-167
  const int sign = dir_sign(dir_bwd, compat);
  std::vector<ocf32> tmp(n);
  for (size_t f = 0; f < howmany; ++f) {
    std::memcpy(tmp.data(), in + f * n, n * sizeof(ocf32));  // tmp[..len].vec_clone(input)
    ora_fft_raw(tmp.data(), out + f * n, n, sign);           // process(tmp, output)
    ora_scale(scale_kind, scale_x, out + f * n, n);          // s.scale(output)
  }
  return ORA_OK;
}

// ----------------------------------------------------------------------------------
// FIR.  src/fir.rs:1-22 defines no filter: y[n] = sum_{k<T} h[k] x[n-k],
// x[<0] = state (T-1 previous inputs, oldest first) or 0.  f32 accumulation in tap
// order; *_f64 is the truth.  frame_len > 0: state is reset to zero at every frame start.
// ----------------------------------------------------------------------------------
int ora_fir(const ocf32* x, size_t n, const ocf32* h, size_t T, ocf32* y, const ocf32* state,
            size_t frame_len) {
  if (T == 0) return ORA_EARG;
  if (frame_len == 0) frame_len = n ? n : 1;
  for (size_t i = 0; i < n; ++i) {
    const size_t f0 = (i / frame_len) * frame_len;  // first sample of this frame
    const size_t in_frame = i - f0 + 1;             // taps that see samples of this frame
    const size_t kmax = T < in_frame ? T : in_frame;
    float ar = 0.0f, ai = 0.0f;
    for (size_t k = 0; k < kmax; ++k) {
      const ocf32 s = x[i - k];
      ar = ar + (h[k].re * s.re - h[k].im * s.im);
      ai = ai + (h[k].re * s.im + h[k].im * s.re);
    }
    if (state && f0 == 0) {
      for (size_t k = kmax; k < T; ++k) {
        // sample at time (i-k) < 0 -> state[(T-1) + (i-k)]
        const long long idx = (long long)(T - 1) + (long long)i - (long long)k;
        if (idx < 0) break;
        const ocf32 s = state[idx];
        ar = ar + (h[k].re * s.re - h[k].im * s.im);
        ai = ai + (h[k].re * s.im + h[k].im * s.re);
      }
    }
    y[i].re = ar; y[i].im = ai;
  }
  return ORA_OK;
}
int ora_fir_f64(const ocf32* x, size_t n, const ocf32* h, size_t T, double* y, const ocf32* state,
                size_t frame_len) {
  if (T == 0) return ORA_EARG;
  if (frame_len == 0) frame_len = n ? n : 1;
  for (size_t i = 0; i < n; ++i) {
    const size_t f0 = (i / frame_len) * frame_len;
    double ar = 0.0, ai = 0.0;
    for (size_t k = 0; k < T; ++k) {
      ocf32 s;
      if (i >= f0 + k) s = x[i - k];
      else if (state && f0 == 0) {
        const long long idx = (long long)(T - 1) + (long long)i - (long long)k;
        if (idx < 0) continue;
        s = state[idx];
      } else continue;
      ar += (double)h[k].re * s.re - (double)h[k].im * s.im;
      ai += (double)h[k].re * s.im + (double)h[k].im * s.re;
    }
    y[2 * i] = ar; y[2 * i + 1] = ai;
  }
  return ORA_OK;
}

// ----------------------------------------------------------------------------------
// sampling  (src/sampling.rs:7-62)
// ----------------------------------------------------------------------------------
// interpolate :7-24.  Writes (n-1)*(k+1)+1 samples to dst (the reference APPENDS them to
// a Vec).  compat=reference keeps `im: x1.re + i*rate.1` (:19).
int ora_interpolate(const ocf32* src, size_t n, ocf32* dst, size_t n_between, int compat) {
  if (n == 0) return ORA_EARG;  // src.last().unwrap() panics :23
  size_t o = 0;
  const float div = (float)(n_between + 1);
  for (size_t w = 0; w + 1 < n; ++w) {
    const ocf32 x1 = src[w], x2 = src[w + 1];
    const float r0 = (x2.re - x1.re) / div;
    const float r1 = (x2.im - x1.im) / div;
    for (size_t ii = 0; ii <= n_between; ++ii) {
      const float i = (float)ii;
      const float t0 = i * r0, t1 = i * r1;
      dst[o].re = x1.re + t0;
      dst[o].im = (compat == ORA_COMPAT_REFERENCE ? x1.re : x1.im) + t1;
      ++o;
    }
  }
  dst[o] = src[n - 1];
  return ORA_OK;
}
// downsample / downsample_sb :28-62 (generic T: Copy -> element size in bytes).
// strict != 0 reproduces the debug_assert (:32,:53) that `cargo test` enforces.
int ora_downsample(const void* src, size_t n_src, void* dst, size_t n_dst, size_t elem, int strict) {
  if (n_dst == 0) return ORA_EARG;  // division by zero panic
  if (strict && (n_src % n_dst) != 0) return ORA_ELEN;
  const size_t dec = n_src / n_dst;
  for (size_t i = 0; i < n_dst; ++i) {
    if (i * dec >= n_src) return ORA_EIDX;
    std::memcpy((char*)dst + i * elem, (const char*)src + i * dec * elem, elem);
  }
  return ORA_OK;
}

// ----------------------------------------------------------------------------------
// modulation  (src/modulation.rs)
// table_len = 2 -> impl Modulation for [cf32;2] (:5-16), 4 -> [cf32;4] (:19-57)
// ----------------------------------------------------------------------------------
// modulate :115-121.  returns ORA_EIDX where Rust would panic on an out-of-bounds index
// (ragged QPSK tail -> bits[1] OOB :24; table index >= table_len :14,:28).
int ora_modulate(const ocf32* table, size_t table_len, const uint8_t* bits, size_t nbits,
                 ocf32* out, size_t out_cap, size_t* n_out) {
  if (table_len != 2 && table_len != 4) return ORA_EARG;
  const size_t bps = table_len == 2 ? 1 : 2;
  size_t o = 0;
  for (size_t i = 0; i < nbits; i += bps) {
    size_t idx;
    if (bps == 1) idx = bits[i];                                     // :10
    else {
      if (i + 1 >= nbits) return ORA_EIDX;                           // bits[1] out of bounds
      idx = (uint8_t)((uint8_t)(bits[i + 1] << 1) + bits[i]);        // :24, u8 arithmetic
    }
    if (idx >= table_len) return ORA_EIDX;
    if (o < out_cap) out[o] = table[idx];                            // modulate_into zip-truncates :123-131
    ++o;
  }
  if (n_out) *n_out = o < out_cap ? o : out_cap;
  return ORA_OK;
}

static inline unsigned argmin_first_wins(const float* d, unsigned cnt) {
  // Iterator::min_by with partial_cmp(..).unwrap_or(Greater): the later element replaces
  // the current best only when compare(best, later) == Greater, i.e. best > later or the
  // two are unordered (NaN).
  unsigned best = 0;
  for (unsigned i = 1; i < cnt; ++i)
    if (!(d[best] <= d[i])) best = i;
  return best;
}

// demod_naive.  table_len 2 -> generic default impl :133-144 (candidates 0..BPS*2);
// table_len 4 -> the QPSK override :33-56, which emits idx&1 then idx&2 in {0,2}
// (SURVEY F5); compat=corrected emits (idx>>1)&1.  Writes BPS bytes per symbol.
int ora_demod(const ocf32* table, size_t table_len, const ocf32* sym, size_t n, uint8_t* bits,
              int compat) {
  if (table_len != 2 && table_len != 4) return ORA_EARG;
  if (table_len == 2) {
    for (size_t i = 0; i < n; ++i) {
      float d[2];
      for (unsigned c = 0; c < 2; ++c) {
        const float dr = sym[i].re - table[c].re, di = sym[i].im - table[c].im;
        d[c] = dr * dr + di * di;
      }
      const unsigned idx = argmin_first_wins(d, 2);
      bits[i] = (uint8_t)(idx & 1u);
    }
  } else {
    for (size_t i = 0; i < n; ++i) {
      float d[4];
      for (unsigned c = 0; c < 4; ++c)
        d[c] = (sym[i].re - table[c].re) * (sym[i].re - table[c].re) +
               (sym[i].im - table[c].im) * (sym[i].im - table[c].im);
      const unsigned idx = argmin_first_wins(d, 4);
      bits[2 * i] = (uint8_t)(idx & 1u);
      bits[2 * i + 1] = compat == ORA_COMPAT_REFERENCE ? (uint8_t)(idx & 2u) : (uint8_t)((idx >> 1) & 1u);
    }
  }
  return ORA_OK;
}

// ----------------------------------------------------------------------------------
// sequence  (src/sequence.rs)
// ----------------------------------------------------------------------------------
int ora_expand(uint64_t seed, size_t len, uint8_t* out) {  // :18-21
  if (len > 64) return ORA_EARG;  // `seed >> i` with i >= 64 overflows (panic in debug)
  for (size_t i = 0; i < len; ++i) out[i] = (uint8_t)((seed >> i) & 1u);
  return ORA_OK;
}
// generate :47-53 with the generator closure restricted to the LFSR family the crate
// documents (:42, test :62): x[n] = (sum_t x[n - back[t]]) % 2.
int ora_mseq_generate(const uint8_t* init, size_t n_init, const uint32_t* back, size_t n_back,
                      size_t len, uint8_t* out) {
  size_t have = n_init < len ? n_init : len;
  // generate() returns init unchanged (even if longer than len) :48; we write min(len, n_init)
  std::memcpy(out, init, have);
  for (size_t n = n_init; n < len; ++n) {
    unsigned s = 0;
    for (size_t t = 0; t < n_back; ++t) {
      if (back[t] == 0 || back[t] > n) return ORA_EIDX;  // seq[n - back] out of bounds / self
      s += out[n - back[t]];
    }
    out[n] = (uint8_t)(s % 2u);
  }
  return ORA_OK;
}

// ----------------------------------------------------------------------------------
// noise  (src/noise.rs) — Philox4x32-10 + Box-Muller replaces ChaCha20 + ziggurat.
// sample i of stream `stream` uses Philox counter (i>>1 lo, i>>1 hi, stream lo, stream hi),
// key = 64-bit seed, words {0,1} for even i and {2,3} for odd i.
// ----------------------------------------------------------------------------------
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
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

static inline void ora_gauss_pair(uint32_t a, uint32_t b, float* z0, float* z1) {
  const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float u2 = fmaf((float)b, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
  const float r = sqrtf(-2.0f * logf(u1));
  const double ang = 2.0 * 3.14159265358979323846 * (double)u2;
  *z0 = r * (float)std::cos(ang);
  *z1 = r * (float)std::sin(ang);
}

// unit-variance complex normal for global sample index i
void ora_awgn_unit(uint64_t seed, uint64_t stream, uint64_t i, float* re, float* im) {
  const uint64_t c = i >> 1;
  const uint32_t ctr[4] = {(uint32_t)c, (uint32_t)(c >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  ora_philox4x32_10(ctr, key, o);
  if (i & 1) ora_gauss_pair(o[2], o[3], re, im);
  else ora_gauss_pair(o[0], o[1], re, im);
}
// Awgn::fill :62-66 via next() :39-44: (N(0,1) as f32) * scale, scale = sqrt(power) :35
void ora_awgn_fill(ocf32* dst, size_t n, float power, uint64_t seed, uint64_t stream, uint64_t offset) {
  const float scale = sqrtf(power);
  for (size_t i = 0; i < n; ++i) {
    float a, b;
    ora_awgn_unit(seed, stream, offset + i, &a, &b);
    dst[i].re = a * scale; dst[i].im = b * scale;
  }
}
// Awgn::apply :53-59: s += next().scale(sc) — the noise is scaled TWICE in the reference
// (SURVEY F5b); compat=corrected scales once.
void ora_awgn_apply(ocf32* sig, size_t n, float power, uint64_t seed, uint64_t stream,
                    uint64_t offset, int compat) {
  const float scale = sqrtf(power);
  for (size_t i = 0; i < n; ++i) {
    float a, b;
    ora_awgn_unit(seed, stream, offset + i, &a, &b);
    float nr = a * scale, ni = b * scale;
    if (compat == ORA_COMPAT_REFERENCE) { nr = nr * scale; ni = ni * scale; }
    sig[i].re = sig[i].re + nr; sig[i].im = sig[i].im + ni;
  }
}

// ----------------------------------------------------------------------------------
// chains
// ----------------------------------------------------------------------------------
// headline chain, one frame at a time: Cfft::fwd (Scale) -> FIR (zero state per frame) ->
// QPSK demod_naive.  sym_out (optional) receives the FIR output symbols.
static void chain_frames(const ocf32* x, size_t n, size_t f_begin, size_t f_end, const ocf32* h,
                         size_t T, const ocf32* table, int scale_kind, float scale_x, int compat,
                         uint8_t* bits, ocf32* sym_out) {
  std::vector<ocf32> X(n), Y(n);
  for (size_t f = f_begin; f < f_end; ++f) {
    ora_cfft_exec(x + f * n, n, X.data(), n, 1, 0, scale_kind, scale_x, compat);
    ora_fir(X.data(), n, h, T, Y.data(), nullptr, n);
    ora_demod(table, 4, Y.data(), n, bits + 2 * f * n, compat);
    if (sym_out) std::memcpy(sym_out + f * n, Y.data(), n * sizeof(ocf32));
  }
}
int ora_chain_fft_fir_demod(const ocf32* x, size_t n, size_t frames, const ocf32* h, size_t T,
                            const ocf32* table, int scale_kind, float scale_x, int compat,
                            uint8_t* bits, ocf32* sym_out, int nthreads) {
  if (n == 0 || T == 0) return ORA_EARG;
  if (nthreads <= 1) {
    chain_frames(x, n, 0, frames, h, T, table, scale_kind, scale_x, compat, bits, sym_out);
    return ORA_OK;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) {
    const size_t b = frames * t / nthreads, e = frames * (t + 1) / nthreads;
    th.emplace_back([=] { chain_frames(x, n, b, e, h, T, table, scale_kind, scale_x, compat, bits, sym_out); });
  }
  for (auto& t : th) t.join();
  return ORA_OK;
}

// modem loop-back (examples/modem.rs:15-32): modulate -> Awgn::apply -> demod_naive,
// plus the bit-error count the example asserts to be zero.
int ora_modem(const ocf32* table, size_t table_len, const uint8_t* bits, size_t nbits, float power,
              uint64_t seed, uint64_t stream, uint64_t offset, int compat, uint8_t* bits_out,
              ocf32* sym_out, uint64_t* bit_errors) {
  const size_t bps = table_len == 2 ? 1 : 2;
  const size_t nsym = nbits / bps;
  std::vector<ocf32> sym(nsym ? nsym : 1);
  size_t n_out = 0;
  int rc = ora_modulate(table, table_len, bits, nbits, sym.data(), nsym, &n_out);
  if (rc) return rc;
  ora_awgn_apply(sym.data(), nsym, power, seed, stream, offset, compat);
  rc = ora_demod(table, table_len, sym.data(), nsym, bits_out, compat);
  if (rc) return rc;
  if (sym_out) std::memcpy(sym_out, sym.data(), nsym * sizeof(ocf32));
  if (bit_errors) {
    uint64_t e = 0;
    for (size_t i = 0; i < nsym * bps; ++i) e += ((bits[i] != 0) != (bits_out[i] != 0));
    *bit_errors = e;
  }
  return ORA_OK;
}

// OFDM-like chain (BASELINE config 5), restated call by call from the reference's pieces:
//   sequence::generate(expand(frame_id+1, 31), |n,s| (s[n-28]+s[n-31])%2, 2N)  (src/sequence.rs:18-53)
//   -> qpsk().modulate (src/modulation.rs:115-121) -> Cfft::bwd(Scale::SN) (src/fft.rs:173-182)
//   -> Awgn::apply semantics (src/noise.rs:53-59) -> Cfft::fwd(Scale::SN) -> demod_naive (:33-56).
// Noise indexing is the fused kernel's: Philox stream = frame id, block = pos mod N/2, words
// {0,1} for the lower half of the frame and {2,3} for the upper half.
// stats = {bit_errors, n_bits, sum|rx-tx|^2, sum|tx|^2} as doubles.
int ora_ofdm_chain(size_t n, size_t frames, uint64_t first_frame, float noise_power, uint64_t seed,
                   int compat, uint8_t* tx_bits, uint8_t* rx_bits, double* stats, ocf32* rx_sym) {
  if (n < 2 || (n & 1)) return ORA_EARG;
  const ocf32 table[4] = {{1.f, 1.f}, {-1.f, 1.f}, {1.f, -1.f}, {-1.f, -1.f}};
  const uint32_t back[2] = {28, 31};
  const float scale = sqrtf(noise_power);
  std::vector<uint8_t> init(31), bits(2 * n), rbits(2 * n);
  std::vector<ocf32> sym(n), t(n), r(n);
  double errs = 0, nb = 0, ep = 0, rp = 0;
  for (size_t f = 0; f < frames; ++f) {
    const uint64_t fid = first_frame + f;
    ora_expand((fid + 1) & 0x7fffffffull, 31, init.data());
    int rc = ora_mseq_generate(init.data(), 31, back, 2, 2 * n, bits.data());
    if (rc) return rc;
    size_t n_out = 0;
    rc = ora_modulate(table, 4, bits.data(), 2 * n, sym.data(), n, &n_out);
    if (rc) return rc;
    ora_cfft_exec(sym.data(), n, t.data(), n, 1, 1, ORA_SCALE_SN, 1.f, compat);
    for (size_t pos = 0; pos < n; ++pos) {
      const uint64_t i = pos < n / 2 ? 2 * (uint64_t)pos : 2 * (uint64_t)(pos - n / 2) + 1;
      float a, b;
      ora_awgn_unit(seed, fid, i, &a, &b);
      float nr = a * scale, ni = b * scale;
      if (compat == ORA_COMPAT_REFERENCE) { nr = nr * scale; ni = ni * scale; }
      t[pos].re = t[pos].re + nr; t[pos].im = t[pos].im + ni;
    }
    ora_cfft_exec(t.data(), n, r.data(), n, 1, 0, ORA_SCALE_SN, 1.f, compat);
    ora_demod(table, 4, r.data(), n, rbits.data(), compat);
    for (size_t i = 0; i < 2 * n; ++i) errs += ((bits[i] != 0) != (rbits[i] != 0));
    nb += 2.0 * n;
    for (size_t i = 0; i < n; ++i) {
      const double dr = (double)r[i].re - sym[i].re, di = (double)r[i].im - sym[i].im;
      ep += dr * dr + di * di;
      rp += (double)sym[i].re * sym[i].re + (double)sym[i].im * sym[i].im;
    }
    if (tx_bits) std::memcpy(tx_bits + 2 * f * n, bits.data(), 2 * n);
    if (rx_bits) std::memcpy(rx_bits + 2 * f * n, rbits.data(), 2 * n);
    if (rx_sym) std::memcpy(rx_sym + f * n, r.data(), n * sizeof(ocf32));
  }
  if (stats) { stats[0] = errs; stats[1] = nb; stats[2] = ep; stats[3] = rp; }
  return ORA_OK;
}

// ----------------------------------------------------------------------------------
// SURVEY 8(f) rows
// ----------------------------------------------------------------------------------
// compute core of util::plot::waterfall (src/util/plot.rs:46-68): zero pad to a multiple of
// fft_len, per chunk vec_rfft(Scale::SN).vec_mirror(), then c.norm() and DB::from (f64 math,
// src/util/mod.rs:26-34).  levels: chunks*fft_len doubles.
int ora_spectrogram(const ocf32* sym, size_t n, size_t fft_len, int use_db, int compat, double* levels) {
  if (fft_len == 0) return ORA_EARG;
  const size_t chunks = (n + fft_len - 1) / fft_len;
  std::vector<ocf32> pad(fft_len), out(fft_len);
  for (size_t c = 0; c < chunks; ++c) {
    for (size_t i = 0; i < fft_len; ++i) {
      const size_t g = c * fft_len + i;
      pad[i] = g < n ? sym[g] : ocf32{0.f, 0.f};
    }
    ora_cfft_exec(pad.data(), fft_len, out.data(), fft_len, 1, 0, ORA_SCALE_SN, 1.f, compat);
    ora_vec_mirror(out.data(), fft_len);
    for (size_t i = 0; i < fft_len; ++i) {
      const float nrm = hypotf(out[i].re, out[i].im);
      levels[c * fft_len + i] = use_db ? 10.0 * std::log10((double)nrm) : (double)nrm;
    }
  }
  return ORA_OK;
}
// benches/benches.rs:410-416: input.vec_rfft(&mut fft, s).vec_mul(&sig).vec_rifft(&mut fft, s), per frame
int ora_correlate(ocf32* data, size_t n, size_t frames, const ocf32* sig, size_t nsig, int scale_kind, float scale_x, int compat) {
  if (nsig != n) return ORA_ELEN;
  std::vector<ocf32> tmp(n);
  for (size_t f = 0; f < frames; ++f) {
    ocf32* p = data + f * n;
    ora_cfft_exec(p, n, tmp.data(), n, 1, 0, scale_kind, scale_x, compat);
    ora_vec_mul(tmp.data(), n, sig, n);
    ora_cfft_exec(tmp.data(), n, p, n, 1, 1, scale_kind, scale_x, compat);
  }
  return ORA_OK;
}

// VecStats (README.md:90-92 TODO; the reference has no definition, so this restates the product's):
// rank cf32 by norm_sqr = re*re + im*im (f32, unfused), f32 by value; first index wins ties, NaN never
// wins, idx = n when nothing is comparable; f64 sums.  out = {n, min_idx, max_idx}, vals = {min, max},
// sums = {sum_re, sum_im, sum_pow}.
static void ora_stats_any(const float* p, size_t n, int cplx, uint64_t* out, float* vals, double* sums) {
  uint64_t imn = n, imx = n;
  float mn = 0.f, mx = 0.f;
  double sre = 0, sim = 0, spw = 0;
  for (size_t i = 0; i < n; ++i) {
    float v;
    if (cplx) {
      const float re = p[2 * i], im = p[2 * i + 1];
      const float a = re * re, b = im * im;
      v = a + b;
      sre += re; sim += im; spw += (double)re * re + (double)im * im;
    } else {
      v = p[i];
      sre += v; spw += (double)v * v;
    }
    if (v == v) {
      if (imn == n || v < mn) { mn = v; imn = i; }
      if (imx == n || v > mx) { mx = v; imx = i; }
    }
  }
  out[0] = n; out[1] = imn; out[2] = imx;
  vals[0] = mn; vals[1] = mx;
  sums[0] = sre; sums[1] = sim; sums[2] = spw;
}
int ora_vec_stats(const ocf32* v, size_t n, uint64_t* out, float* vals, double* sums) {
  if (n == 0) return ORA_ELEN;
  ora_stats_any((const float*)v, n, 1, out, vals, sums);
  return ORA_OK;
}
int ora_f32_stats(const float* v, size_t n, uint64_t* out, float* vals, double* sums) {
  if (n == 0) return ORA_ELEN;
  ora_stats_any(v, n, 0, out, vals, sums);
  return ORA_OK;
}

}  // extern "C"
