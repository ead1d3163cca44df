// chain_x2_host.h — host-side tables of K14b (chain_x2.cuh): per-thread twiddle rows and the TF32 hi/lo
// split of the taps.  Plain C++ so that the host emulation harness (tests/cpp/chain_x2_emu.cpp) builds
// them with the same code as the library.
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

#include "chain_x2.cuh"

namespace ae {

inline size_t chain_x2_twiddle_count(size_t nfft) { return nfft == 1024 ? (size_t)X2Cfg<1024>::ROWS * X2Cfg<1024>::T : 0; }

// rows of T per-thread values exp(-2 pi i e/len), f64-evaluated (layout: X2Cfg)
inline void chain_x2_twiddles(size_t nfft, std::vector<float2>& out) {
  out.clear();
  if (nfft != 1024) return;
  using XC = X2Cfg<1024>;
  constexpr int N = 1024, T = XC::T;
  out.assign((size_t)XC::ROWS * T, make_float2(1.f, 0.f));
  auto W = [](long long e, long long len) {
    const double a = -2.0 * 3.14159265358979323846 * (double)(e % len) / (double)len;
    return make_float2((float)std::cos(a), (float)std::sin(a));
  };
  for (int t = 0; t < T; ++t) {
    for (int m = 0; m < 16; ++m) out[(size_t)(XC::ROW_S1 + m) * T + t] = W(t + (long long)T * m, N);
    const int mult[6] = {1, 2, 3, 4, 8, 12};   // Z[2 k0 + b] = W_512^(k0 t) Y[2 k0 + b]; k0 = 4a + b' from W^(4a t) W^(b' t)
    for (int e = 0; e < 6; ++e) out[(size_t)(XC::ROW_P1 + e) * T + t] = W((long long)mult[e] * t, N / 2);
  }
}

// taps split for the 3xTF32 products: hi = the tap truncated to TF32's 10 mantissa bits, lo = tap - hi
// (exact in FP32; the tensor core uses its leading 11 bits)
inline void chain_x2_split_taps(const float2* taps, size_t ntaps, std::vector<float2>& hi, std::vector<float2>& lo) {
  const size_t pad = X2Cfg<1024>::HPAD;
  hi.assign(pad, make_float2(0.f, 0.f));
  lo.assign(pad, make_float2(0.f, 0.f));
  auto split = [](float v, float& h, float& l) {
    uint32_t b;
    std::memcpy(&b, &v, 4);
    b &= 0xffffe000u;
    std::memcpy(&h, &b, 4);
    l = v - h;
    if (!(l == l)) l = 0.0f;  // inf - inf
  };
  for (size_t k = 0; k < ntaps && k < pad; ++k) {
    split(taps[k].x, hi[k].x, lo[k].x);
    split(taps[k].y, hi[k].y, lo[k].y);
  }
}

}  // namespace ae
