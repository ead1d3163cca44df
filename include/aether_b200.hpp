// aether_b200.hpp — C++17 host-side mirror of the reference crate's API over the C ABI.
//
// The reference's host language is Rust; no Rust toolchain exists in the build image, so the host
// side above the C ABI is written in C++ (the reference is compiled code) with the crate's own
// names, argument meaning and failure behaviour:
//     aether::DeviceVec        <->  Vec<cf32> / &mut [cf32] implementing `VecOps`  (src/vecops.rs:39-89)
//     aether::Scale, Cfft      <->  fft::Scale, fft::Fft, fft::Cfft                  (src/fft.rs:6-235)
//     aether::Modulation       <->  modulation::Modulation, bpsk(), qpsk()           (src/modulation.rs)
//     aether::Awgn             <->  noise::Awgn, generator(), new()                  (src/noise.rs)
//     aether::sampling::*      <->  sampling::{interpolate, downsample, downsample_sb}
//     aether::sequence::*      <->  sequence::{expand, generate}
// A Rust panic becomes an aether::Panic exception carrying the reference's message.
// Header-only; link with libaether_b200.so.  (INTEGRATION.md shows the Rust binding.)
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "aether_b200.h"

namespace aether {

using cf32 = std::complex<float>;  // src/lib.rs:12 — repr(C) {re, im}
static_assert(sizeof(cf32) == sizeof(ae_cf32), "cf32 layout");

struct Panic : std::runtime_error {
  int status;
  Panic(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};
inline void check(ae_status st) {
  if (st != AE_OK) throw Panic(st, ae_last_error_string());
}
inline void init(int device = 0) { check(ae_init(device)); }
inline void sync() { check(ae_sync()); }

enum class Compat : int { Reference = AE_COMPAT_REFERENCE, Corrected = AE_COMPAT_CORRECTED };

class Cfft;

/// fft::Scale (src/fft.rs:6-18)
struct Scale {
  int kind;
  float x;
  static Scale None() { return {AE_SCALE_NONE, 1.f}; }
  static Scale SN() { return {AE_SCALE_SN, 1.f}; }
  static Scale N() { return {AE_SCALE_N, 1.f}; }
  static Scale X(float v) { return {AE_SCALE_X, v}; }
};

/// device Vec<u8>, one byte per bit (src/modulation.rs:102-103)
class DeviceBits {
 public:
  explicit DeviceBits(size_t capacity = 1) { check(ae_bits_alloc(0, capacity, &h_)); }
  explicit DeviceBits(const std::vector<uint8_t>& host) {
    check(ae_bits_alloc(host.size(), host.size(), &h_));
    check(ae_bits_upload(h_, host.data(), host.size()));
  }
  DeviceBits(const DeviceBits&) = delete;
  DeviceBits& operator=(const DeviceBits&) = delete;
  DeviceBits(DeviceBits&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~DeviceBits() { ae_bits_free(h_); }
  size_t len() const { return ae_bits_len(h_); }
  void clear() { check(ae_bits_set_len(h_, 0)); }
  std::vector<uint8_t> to_host() const {
    std::vector<uint8_t> v(len());
    check(ae_bits_download(h_, v.data(), v.size()));
    return v;
  }
  ae_bits* raw() const { return h_; }

 private:
  ae_bits* h_ = nullptr;
};

/// device Vec<cf32> implementing the reference's VecOps trait: every method returns *this so
/// calls chain exactly like `v.vec_div(&a).vec_mul(&b)...` (src/vecops.rs:19-36); the chain is
/// recorded and runs as ONE fused kernel when the data is needed.
class DeviceVec {
 public:
  explicit DeviceVec(size_t len) { check(ae_vec_alloc(len, len, &h_)); }
  static DeviceVec with_capacity(size_t cap) {
    DeviceVec v;
    check(ae_vec_alloc(0, cap, &v.h_));
    return v;
  }
  explicit DeviceVec(const std::vector<cf32>& host) {
    check(ae_vec_alloc(host.size(), host.size(), &h_));
    check(ae_vec_upload(h_, reinterpret_cast<const ae_cf32*>(host.data()), host.size()));
  }
  DeviceVec(const DeviceVec&) = delete;
  DeviceVec& operator=(const DeviceVec&) = delete;
  DeviceVec(DeviceVec&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~DeviceVec() { ae_vec_free(h_); }

  size_t len() const { return ae_vec_len(h_); }
  size_t capacity() const { return ae_vec_capacity(h_); }
  std::vector<cf32> to_host() const {
    std::vector<cf32> v(len());
    check(ae_vec_download(h_, reinterpret_cast<ae_cf32*>(v.data()), v.size()));
    return v;
  }
  /// &mut v[start..end]
  DeviceVec slice(size_t start, size_t end) {
    DeviceVec v;
    check(ae_vec_view(h_, start, end - start, &v.h_));
    return v;
  }
  // ---- VecOps (src/vecops.rs:39-89) ----
  DeviceVec& vec_scale(float s) { check(ae_vec_scale(h_, s)); return *this; }
  DeviceVec& vec_mul(const DeviceVec& o) { check(ae_vec_mul(h_, o.h_)); return *this; }
  DeviceVec& vec_div(const DeviceVec& o) { check(ae_vec_div(h_, o.h_)); return *this; }
  DeviceVec& vec_conj() { check(ae_vec_conj(h_)); return *this; }
  DeviceVec& vec_mirror() { check(ae_vec_mirror(h_)); return *this; }
  DeviceVec& vec_clone(const DeviceVec& o) { check(ae_vec_clone(h_, o.h_)); return *this; }
  DeviceVec& vec_zero() { check(ae_vec_zero(h_)); return *this; }
  DeviceVec& vec_add(const DeviceVec& o) { check(ae_vec_add(h_, o.h_)); return *this; }
  DeviceVec& vec_sub(const DeviceVec& o) { check(ae_vec_sub(h_, o.h_)); return *this; }
  /// arbitrary closure in element order: host round trip (documented slow path)
  template <class F>
  DeviceVec& vec_mutate(F&& f) {
    auto tramp = [](ae_cf32* e, void* u) { (*static_cast<F*>(u))(*reinterpret_cast<cf32*>(e)); };
    check(ae_vec_mutate(h_, tramp, &f));
    return *this;
  }
  DeviceVec& vec_fft(Scale s, Compat c = Compat::Reference) { check(ae_vec_fft(h_, s.kind, s.x, (int)c)); return *this; }
  DeviceVec& vec_ifft(Scale s, Compat c = Compat::Reference) { check(ae_vec_ifft(h_, s.kind, s.x, (int)c)); return *this; }
  inline DeviceVec& vec_rfft(Cfft& fft, Scale s);
  inline DeviceVec& vec_rifft(Cfft& fft, Scale s);
  /// VecStats (README.md:90-92 TODO): min/max by norm_sqr with their indices, f64 sums; mean = sum/n
  ae_vecstats vec_stats() const {
    ae_vecstats st;
    check(ae_vec_stats(h_, &st));
    return st;
  }
  ae_vec* raw() const { return h_; }

 private:
  DeviceVec() = default;
  friend class Cfft;
  ae_vec* h_ = nullptr;
};

/// fft::Cfft implementing fft::Fft (src/fft.rs:48-77, :134-235); `howmany` frames per call is new surface
class Cfft {
 public:
  static Cfft with_len(size_t len, Compat c = Compat::Reference) { return Cfft(len, c); }
  Cfft(size_t len, Compat c = Compat::Reference) {
    check(ae_fft_create(len, &h_));
    check(ae_fft_set_compat(h_, (int)c));
  }
  Cfft(const Cfft&) = delete;
  Cfft(Cfft&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~Cfft() { ae_fft_destroy(h_); }
  size_t len() const { return ae_fft_len(h_); }
  void fwd(const DeviceVec& in, DeviceVec& out, Scale s, size_t howmany = 1) { check(ae_fft_exec(h_, AE_FFT_FWD, in.raw(), out.raw(), s.kind, s.x, howmany)); }
  void bwd(const DeviceVec& in, DeviceVec& out, Scale s, size_t howmany = 1) { check(ae_fft_exec(h_, AE_FFT_BWD, in.raw(), out.raw(), s.kind, s.x, howmany)); }
  void ifwd(DeviceVec& io, Scale s, size_t howmany = 1) { check(ae_fft_exec(h_, AE_FFT_FWD, io.raw(), nullptr, s.kind, s.x, howmany)); }
  void ibwd(DeviceVec& io, Scale s, size_t howmany = 1) { check(ae_fft_exec(h_, AE_FFT_BWD, io.raw(), nullptr, s.kind, s.x, howmany)); }
  /// tfwd/tbwd: the returned handle borrows the plan's scratch until the next call on this plan
  ae_vec* tfwd(const DeviceVec& in, Scale s, size_t howmany = 1) { ae_vec* v; check(ae_fft_exec_tmp(h_, AE_FFT_FWD, in.raw(), s.kind, s.x, howmany, &v)); return v; }
  ae_vec* tbwd(const DeviceVec& in, Scale s, size_t howmany = 1) { ae_vec* v; check(ae_fft_exec_tmp(h_, AE_FFT_BWD, in.raw(), s.kind, s.x, howmany, &v)); return v; }

 private:
  ae_fft* h_ = nullptr;
};
inline DeviceVec& DeviceVec::vec_rfft(Cfft& fft, Scale s) { fft.ifwd(*this, s); return *this; }
inline DeviceVec& DeviceVec::vec_rifft(Cfft& fft, Scale s) { fft.ibwd(*this, s); return *this; }

/// the filter src/fir.rs:1-22 only sketches
class Fir {
 public:
  Fir(const std::vector<cf32>& taps, size_t /*input_len*/ = 0, int mode = AE_FIR_AUTO) {
    check(ae_fir_create(reinterpret_cast<const ae_cf32*>(taps.data()), taps.size(), mode, &h_));
  }
  Fir(const Fir&) = delete;
  ~Fir() { ae_fir_destroy(h_); }
  void reset() { check(ae_fir_reset(h_)); }
  void filter(const DeviceVec& in, DeviceVec& out, size_t frame_len = 0) { check(ae_fir_exec(h_, in.raw(), out.raw(), frame_len)); }

 private:
  ae_fir* h_ = nullptr;
};

/// impl Modulation for [cf32; 2] / [cf32; 4] (src/modulation.rs:5-57, :94-149)
class Modulation {
 public:
  explicit Modulation(const std::vector<cf32>& table) { check(ae_mod_create(reinterpret_cast<const ae_cf32*>(table.data()), table.size(), &h_)); }
  Modulation(const Modulation&) = delete;
  Modulation(Modulation&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~Modulation() { ae_mod_destroy(h_); }
  size_t bits_per_symbol() const { return ae_mod_bits_per_symbol(h_); }
  DeviceVec modulate(const DeviceBits& bits) {
    DeviceVec out = DeviceVec::with_capacity(bits.len() / bits_per_symbol() + 1);
    check(ae_mod_modulate(h_, bits.raw(), out.raw()));
    return out;
  }
  void modulate_into(const DeviceBits& bits, DeviceVec& out) { check(ae_mod_modulate_into(h_, bits.raw(), out.raw())); }
  void demod_naive(const DeviceVec& symbols, DeviceBits& out, Compat c = Compat::Reference) { check(ae_mod_demod(h_, symbols.raw(), out.raw(), (int)c)); }
  ae_mod* raw() const { return h_; }

 private:
  ae_mod* h_ = nullptr;
};
inline Modulation bpsk() { return Modulation({{1.f, 1.f}, {-1.f, -1.f}}); }                              // src/modulation.rs:61-63, :77
inline Modulation qpsk() { return Modulation({{1.f, 1.f}, {-1.f, 1.f}, {1.f, -1.f}, {-1.f, -1.f}}); }    // :66-68, :87-92

/// noise::Awgn (src/noise.rs:20-71)
class Awgn {
 public:
  Awgn(float power, uint64_t seed) { check(ae_awgn_create(power, seed, &h_)); }
  Awgn(const Awgn&) = delete;
  Awgn(Awgn&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~Awgn() { ae_awgn_destroy(h_); }
  void set_power(float p) { check(ae_awgn_set_power(h_, p)); }
  void apply(DeviceVec& signal, Compat c = Compat::Reference) { check(ae_awgn_apply(h_, signal.raw(), (int)c)); }
  void fill(DeviceVec& target) { check(ae_awgn_fill(h_, target.raw())); }
  ae_awgn* raw() const { return h_; }

 private:
  ae_awgn* h_ = nullptr;
};
namespace noise {
inline Awgn generator() { return Awgn(1.f, 815); }                    // src/noise.rs:9-11
inline Awgn new_(float power, uint64_t seed) { return Awgn(power, seed); }  // src/noise.rs:14-16
}  // namespace noise

namespace sampling {  // src/sampling.rs
inline void interpolate(const DeviceVec& src, DeviceVec& dst, size_t n_between, Compat c = Compat::Reference) { check(ae_interpolate(src.raw(), dst.raw(), n_between, (int)c)); }
inline void downsample(const DeviceVec& src, DeviceVec& dst, bool strict = true) { check(ae_downsample(src.raw(), dst.raw(), strict)); }
inline void downsample_sb(const DeviceVec& src, DeviceVec& dst, bool strict = true) { check(ae_downsample_sb(src.raw(), dst.raw(), strict)); }
}  // namespace sampling

namespace sequence {  // src/sequence.rs
inline DeviceBits expand(uint64_t seed, size_t len) {
  DeviceBits b(len ? len : 1);
  check(ae_mseq_expand(seed, len, b.raw()));
  return b;
}
/// generate() with the closure given as its tap list: x[n] = (sum_t x[n - back[t]]) % 2
inline DeviceBits generate(const std::vector<uint8_t>& init, const std::vector<uint32_t>& back, size_t len) {
  DeviceBits b(len > init.size() ? len : init.size() + 1);
  check(ae_mseq_generate(init.data(), init.size(), back.data(), back.size(), len, b.raw()));
  return b;
}
}  // namespace sequence

/// headline chain: per frame Cfft::fwd(scale) -> FIR (zero state per frame) -> QPSK demod_naive, one kernel
class FftFirDemod {
 public:
  FftFirDemod(size_t fft_len, const std::vector<cf32>& taps, Scale s, Compat c = Compat::Reference) {
    check(ae_chain_create(fft_len, reinterpret_cast<const ae_cf32*>(taps.data()), taps.size(), s.kind, s.x, (int)c, &h_));
  }
  FftFirDemod(const FftFirDemod&) = delete;
  ~FftFirDemod() { ae_chain_destroy(h_); }
  void run(const DeviceVec& in, DeviceBits& bits_out) { check(ae_chain_exec(h_, in.raw(), bits_out.raw())); }
  /// host buffers (ideally from ae_host_alloc): chunked H2D -> kernel -> D2H inside the call
  void run_host(const cf32* host_in, size_t n_samples, uint8_t* host_bits) {
    check(ae_chain_exec_host(h_, reinterpret_cast<const ae_cf32*>(host_in), n_samples, host_bits));
  }
  ae_chain* raw() const { return h_; }

 private:
  ae_chain* h_ = nullptr;
};

/// pipeline::Pipeline / pool::Pool analogue for this path (src/pipeline.rs:26-137, src/pool.rs:43-130):
/// send() queues a block on one of `depth` buffer slots, recv() returns finished bit buffers in order,
/// report() is the per-stage line the reference prints (processed, active time, rate, utilisation).
class ChainPipeline {
 public:
  ChainPipeline(FftFirDemod& chain, size_t block_frames, int depth = 3) { check(ae_pipe_create(chain.raw(), block_frames, depth, &h_)); }
  ChainPipeline(const ChainPipeline&) = delete;
  ~ChainPipeline() { ae_pipe_destroy(h_); }
  void send(const cf32* host_in, uint8_t* host_bits) { check(ae_pipe_send(h_, reinterpret_cast<const ae_cf32*>(host_in), host_bits)); }
  uint8_t* recv() {
    uint8_t* p = nullptr;
    check(ae_pipe_recv(h_, &p));
    return p;
  }
  size_t in_flight() const { return ae_pipe_in_flight(h_); }
  std::vector<ae_pipe_stage> report(bool reset = false) {
    std::vector<ae_pipe_stage> st(3);
    check(ae_pipe_report(h_, st.data(), reset));
    return st;
  }

 private:
  ae_pipe* h_ = nullptr;
};

/// util::DB (src/util/mod.rs:11-46): a value in decibel; DB::from(ratio) = 10 log10(ratio) in f64
struct DB {
  double value;
  static DB from(double ratio) { return DB{10.0 * std::log10(ratio)}; }
  double db() const { return value; }
  double ratio() const { return std::pow(10.0, value / 10.0); }
};

/// assert_evm!(actual, ref[, evm_limit_db]) (src/lib.rs:26-49) on host data: per element
/// |act - re| > |re| * (10^(dB/10) as f32) panics (here: throws Panic with the macro's message text).
inline void assert_evm(const std::vector<cf32>& actual, const std::vector<cf32>& ref, double evm_limit_db = -80.0) {
  if (actual.size() != ref.size()) throw Panic(AE_ELEN, "Input slices/vectors must be same length");
  if (!(evm_limit_db < 0.0)) throw Panic(AE_EARG, "The EVM threshold must be negative");
  const float factor = (float)std::pow(10.0, evm_limit_db / 10.0);
  for (size_t i = 0; i < actual.size(); ++i) {
    const float evm = std::abs(actual[i] - ref[i]);
    const float limit = std::abs(ref[i]) * factor;
    if (evm > limit) throw Panic(AE_EARG, "EVM limit exceeded for element " + std::to_string(i));
  }
}

}  // namespace aether
