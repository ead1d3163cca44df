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
//     aether::DeviceStats, Comm, ShardedFir, Graph, chain::*  — the multi-GPU and fused surface of this path
//     (BER/EVM counters + their NCCL reduction, stream sharding with a filter halo, CUDA-graph replay,
//      modem / OFDM / spectrogram / correlator kernels); tests/test_abi_cpu.py checks that every ae_* entry point of
//      the C header is reachable from this file.
// A Rust panic becomes an aether::Panic exception carrying the reference's message.
// Header-only; link with libaether_b200.so.  (INTEGRATION.md shows the Rust binding.)
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <optional>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
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
inline std::string version() { return ae_version(); }
inline int device_count() { int n = 0; check(ae_device_count(&n)); return n; }
inline int sm_count() { int n = 0; check(ae_sm_count(&n)); return n; }
/// run on a caller-owned cudaStream_t (nullptr: the library's own stream)
inline void set_stream(void* cuda_stream) { check(ae_set_stream(cuda_stream)); }
inline void* get_stream() { return ae_get_stream(); }
/// kernels launched by the library so far (bench.py's gpu_launches)
inline uint64_t launch_count() { return ae_launch_count(); }

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
  /// Scale::scale(len) (src/fft.rs:21-45): the f32 factor the kernels multiply with
  float factor(size_t n) const { float f = 0.f; check(ae_scale_factor(kind, n, x, &f)); return f; }
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
  /// borrow foreign device memory (the caller keeps it alive)
  static DeviceBits wrap(void* device_ptr, size_t len) {
    DeviceBits b(nullptr);
    check(ae_bits_wrap(device_ptr, len, &b.h_));
    return b;
  }
  size_t len() const { return ae_bits_len(h_); }
  size_t capacity() const { return ae_bits_capacity(h_); }
  void* device_ptr() const { void* q = nullptr; check(ae_bits_device_ptr(h_, &q)); return q; }
  void clear() { check(ae_bits_set_len(h_, 0)); }
  void upload_async(const uint8_t* host, size_t n) { check(ae_bits_upload_async(h_, host, n)); }
  void download_async(uint8_t* host, size_t n) const { check(ae_bits_download_async(h_, host, n)); }
  std::vector<uint8_t> to_host() const {
    std::vector<uint8_t> v(len());
    check(ae_bits_download(h_, v.data(), v.size()));
    return v;
  }
  ae_bits* raw() const { return h_; }

 private:
  explicit DeviceBits(std::nullptr_t) {}
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

  /// borrow foreign device memory (the caller keeps it alive)
  static DeviceVec wrap(void* device_ptr, size_t len) {
    DeviceVec v;
    check(ae_vec_wrap(device_ptr, len, &v.h_));
    return v;
  }
  /// raw interleaved f32 (re, im) sample file, the format of the crate's file sources/sinks
  static DeviceVec read_raw(const std::string& path) {
    DeviceVec v;
    check(ae_vec_read_raw(path.c_str(), &v.h_));
    return v;
  }
  void write_raw(const std::string& path) { check(ae_vec_write_raw(h_, path.c_str())); }
  size_t len() const { return ae_vec_len(h_); }
  size_t capacity() const { return ae_vec_capacity(h_); }
  void reserve(size_t cap) { check(ae_vec_reserve(h_, cap)); }     // Vec::reserve
  void set_len(size_t n) { check(ae_vec_set_len(h_, n)); }         // Vec::truncate / unsafe set_len within capacity
  void clear() { set_len(0); }
  void* device_ptr() const { void* q = nullptr; check(ae_vec_device_ptr(h_, &q)); return q; }
  /// the recorded VecOps chain runs when the data is needed; flush() runs it now
  size_t pending_ops() const { return ae_vec_pending_ops(h_); }
  void flush() { check(ae_vec_flush(h_)); }
  /// stream-ordered copies for pipeline stages: no synchronisation, host memory pinned and valid until the item is received
  void upload_async(const cf32* host, size_t n) { check(ae_vec_upload_async(h_, reinterpret_cast<const ae_cf32*>(host), n)); }
  void download_async(cf32* host, size_t n) const { check(ae_vec_download_async(h_, reinterpret_cast<ae_cf32*>(host), n)); }
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
  DeviceVec& vec_scale(Scale s) { check(ae_vec_scale_kind(h_, s.kind, s.x)); return *this; }   // v.vec_scale(Scale::SN.scale(len))
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
  friend class ShardedFir;
  ae_vec* h_ = nullptr;
};

/// device Vec<f32> (spectrogram levels)
class DeviceF32 {
 public:
  explicit DeviceF32(size_t len) { check(ae_f32_alloc(len, &h_)); }
  DeviceF32(const DeviceF32&) = delete;
  DeviceF32(DeviceF32&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~DeviceF32() { ae_f32_free(h_); }
  size_t len() const { return ae_f32_len(h_); }
  void* device_ptr() const { void* q = nullptr; check(ae_f32_device_ptr(h_, &q)); return q; }
  std::vector<float> to_host() const {
    std::vector<float> v(len());
    check(ae_f32_download(h_, v.data(), v.size()));
    return v;
  }
  ae_vecstats stats() const { ae_vecstats st; check(ae_f32_stats(h_, &st)); return st; }
  ae_f32* raw() const { return h_; }

 private:
  ae_f32* h_ = nullptr;
};

/// pinned host memory for the *_host entry points and the pipeline
template <class T>
class PinnedBuf {
 public:
  explicit PinnedBuf(size_t n) : n_(n) { void* q = nullptr; check(ae_host_alloc(n * sizeof(T), &q)); p_ = static_cast<T*>(q); }
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf(PinnedBuf&& o) noexcept : p_(std::exchange(o.p_, nullptr)), n_(std::exchange(o.n_, 0)) {}
  ~PinnedBuf() { if (p_) ae_host_free(p_); }
  T* data() { return p_; }
  const T* data() const { return p_; }
  size_t size() const { return n_; }
  T& operator[](size_t i) { return p_[i]; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
};

/// {bit_errors, n_bits, sum|e|^2, sum|r|^2} on the device: what examples/modem.rs counts on the host
class DeviceStats {
 public:
  DeviceStats() { check(ae_stats_alloc(&h_)); }
  DeviceStats(const DeviceStats&) = delete;
  ~DeviceStats() { ae_stats_free(h_); }
  void zero() { check(ae_stats_zero(h_)); }
  ae_stats read() const { ae_stats s; check(ae_stats_read(h_, &s)); return s; }
  double ber() const { const ae_stats s = read(); return s.n_bits ? (double)s.bit_errors / (double)s.n_bits : 0.0; }
  double evm_db() const { const ae_stats s = read(); return 10.0 * std::log10(s.err_pow / s.ref_pow); }
  void count_bit_errors(const DeviceBits& a, const DeviceBits& b) { check(ae_count_bit_errors(a.raw(), b.raw(), h_)); }
  void evm_accumulate(const DeviceVec& actual, const DeviceVec& reference) { check(ae_evm_accumulate(actual.raw(), reference.raw(), h_)); }
  ae_stats* raw() const { return h_; }

 private:
  ae_stats* h_ = nullptr;
};

/// NCCL communicator behind the C ABI (the library binds libnccl at run time).  One process per GPU: rank 0 makes the
/// id, every rank calls init_rank on its own device; one process, n GPUs: init_all + allreduce_all.
class Comm {
 public:
  static std::vector<uint8_t> unique_id() {
    std::vector<uint8_t> id(AE_COMM_ID_BYTES);
    check(ae_comm_unique_id(id.data()));
    return id;
  }
  static Comm init_rank(const std::vector<uint8_t>& id, int nranks, int rank) {
    if (id.size() != AE_COMM_ID_BYTES) throw Panic(AE_EARG, "communicator id must be AE_COMM_ID_BYTES long");
    Comm c;
    check(ae_comm_init_rank(id.data(), nranks, rank, &c.h_));
    return c;
  }
  static std::vector<Comm> init_all(int ndev) {
    std::vector<ae_comm*> raw((size_t)ndev, nullptr);
    check(ae_comm_init_all(ndev, raw.data()));
    std::vector<Comm> out;
    for (ae_comm* r : raw) { Comm c; c.h_ = r; out.push_back(std::move(c)); }
    return out;
  }
  Comm(const Comm&) = delete;
  Comm(Comm&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~Comm() { if (h_) ae_comm_destroy(h_); }
  struct Info { int nranks, rank, device; };
  Info info() const { Info i{0, 0, 0}; check(ae_comm_info(h_, &i.nranks, &i.rank, &i.device)); return i; }
  /// sum the counters over all ranks, in place on the device, on the context stream
  void allreduce(DeviceStats& st) { check(ae_stats_allreduce(st.raw(), h_)); }
  static void allreduce_all(std::vector<DeviceStats*>& stats, std::vector<Comm>& comms) {
    std::vector<ae_stats*> s;
    std::vector<ae_comm*> c;
    for (auto* x : stats) s.push_back(x->raw());
    for (auto& x : comms) c.push_back(x.h_);
    check(ae_stats_allreduce_all(s.data(), c.data(), (int)c.size()));
  }
  ae_comm* raw() const { return h_; }

 private:
  Comm() = default;
  ae_comm* h_ = nullptr;
};

/// CUDA-graph replay of everything the library launches between begin and end (launch-bound small steps)
class Graph {
 public:
  static void begin() { check(ae_graph_begin()); }
  static Graph end() { Graph g; check(ae_graph_end(&g.h_)); return g; }
  template <class F>
  static Graph record(F&& body) { begin(); body(); return end(); }
  Graph(const Graph&) = delete;
  Graph(Graph&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  ~Graph() { if (h_) ae_graph_destroy(h_); }
  void launch() { check(ae_graph_launch(h_)); }

 private:
  Graph() = default;
  ae_graph* h_ = nullptr;
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
  /// per frame: |fwd(frame)|^2 as power or dB, DC in the middle (the crate's waterfall / spectrum plots)
  void spectrogram(const DeviceVec& symbols, DeviceF32& levels, bool use_db) { check(ae_spectrogram(h_, symbols.raw(), levels.raw(), use_db)); }
  /// per frame: inout <- bwd(fwd(inout) * conj(fwd(sig))): the frequency-domain correlator
  void correlate(DeviceVec& inout, const DeviceVec& sig, Scale s, size_t howmany = 1) { check(ae_correlate(h_, inout.raw(), sig.raw(), s.kind, s.x, howmany)); }
  ae_fft* raw() const { return h_; }

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
  size_t ntaps() const { return ae_fir_ntaps(h_); }
  /// outputs per overlap-save block (1 for the direct form): shard boundaries on a multiple of it reproduce the
  /// unsharded stream bit for bit
  size_t block_hop() const { return ae_fir_block_hop(h_); }

 private:
  ae_fir* h_ = nullptr;
};

/// impl Modulation for [cf32; 2] / [cf32; 4] (src/modulation.rs:5-57, :94-149)
class Modulation {
 public:
  explicit Modulation(const std::vector<cf32>& table) { check(ae_mod_create(reinterpret_cast<const ae_cf32*>(table.data()), table.size(), &h_)); }
  explicit Modulation(ae_mod* adopted) : h_(adopted) {}
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
inline Modulation bpsk() { ae_mod* m = nullptr; check(ae_mod_bpsk(&m)); return Modulation(m); }   // src/modulation.rs:61-63, :77
inline Modulation qpsk() { ae_mod* m = nullptr; check(ae_mod_qpsk(&m)); return Modulation(m); }   // :66-68, :87-92

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
  /// Iterator::next() n times, on the host (the same stream the device kernels draw from)
  std::vector<cf32> next(size_t n) {
    std::vector<cf32> v(n);
    check(ae_awgn_next_host(h_, reinterpret_cast<ae_cf32*>(v.data()), n));
    return v;
  }
  /// position in the stream, in samples (results depend only on (seed, stream id, sample index))
  uint64_t tell() const { return ae_awgn_tell(h_); }
  void seek(uint64_t sample_offset) { check(ae_awgn_seek(h_, sample_offset)); }
  void set_stream_id(uint64_t id) { check(ae_awgn_set_stream_id(h_, id)); }     // one independent stream per rank / shard
  explicit Awgn(ae_awgn* adopted) : h_(adopted) {}
  ae_awgn* raw() const { return h_; }

 private:
  ae_awgn* h_ = nullptr;
};
namespace noise {
inline Awgn generator() { ae_awgn* g = nullptr; check(ae_awgn_generator(&g)); return Awgn(g); }   // src/noise.rs:9-11
inline Awgn new_(float power, uint64_t seed) { return Awgn(power, seed); }  // src/noise.rs:14-16
/// one Philox4x32-10 block (the generator under Awgn; known-answer vectors in tests/test_noise.py)
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { check(ae_philox4x32_10(ctr, key, out)); }
}  // namespace noise

namespace sampling {  // src/sampling.rs
inline void interpolate(const DeviceVec& src, DeviceVec& dst, size_t n_between, Compat c = Compat::Reference) { check(ae_interpolate(src.raw(), dst.raw(), n_between, (int)c)); }
inline void downsample(const DeviceVec& src, DeviceVec& dst, bool strict = true) { check(ae_downsample(src.raw(), dst.raw(), strict)); }
inline void downsample_sb(const DeviceVec& src, DeviceVec& dst, bool strict = true) { check(ae_downsample_sb(src.raw(), dst.raw(), strict)); }
inline void downsample(const DeviceBits& src, DeviceBits& dst, bool strict = true) { check(ae_downsample_bits(src.raw(), dst.raw(), strict)); }   // generic over T
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
  /// the same chain as the composition of the stand-alone kernels (parity reference for the fused one)
  void run_unfused(const DeviceVec& in, DeviceBits& bits_out, DeviceVec* symbols_out = nullptr) {
    check(ae_chain_exec_unfused(h_, in.raw(), bits_out.raw(), symbols_out ? symbols_out->raw() : nullptr));
  }
  ae_chain* raw() const { return h_; }

 private:
  ae_chain* h_ = nullptr;
};

namespace chain {
/// examples/modem.rs:15-32 in one kernel: modulate -> Awgn::apply -> demod_naive, bit errors counted on the device
inline void modem_fused(Modulation& m, Awgn& g, const DeviceBits& bits_in, DeviceBits& bits_out, DeviceStats* stats = nullptr,
                        Compat c = Compat::Reference) {
  check(ae_modem_fused(m.raw(), g.raw(), bits_in.raw(), bits_out.raw(), stats ? stats->raw() : nullptr, (int)c));
}
/// M-sequence -> QPSK -> bwd FFT(SN) -> AWGN -> fwd FFT(SN) -> demod -> BER/EVM counters, one kernel; frames are
/// numbered globally (first_frame_id) so N ranks produce the counters of one rank doing all of them
inline void ofdm_chain(size_t fft_len, size_t frames, uint64_t first_frame_id, float noise_power, uint64_t noise_seed, DeviceStats* stats,
                       DeviceBits* tx_bits = nullptr, DeviceBits* rx_bits = nullptr, Compat c = Compat::Reference) {
  check(ae_ofdm_chain(fft_len, frames, first_frame_id, noise_power, noise_seed, (int)c, tx_bits ? tx_bits->raw() : nullptr,
                      rx_bits ? rx_bits->raw() : nullptr, stats ? stats->raw() : nullptr));
}
}  // namespace chain

/// One continuous stream filtered by `world` ranks (BASELINE config 3): rank r owns outputs [lo, hi) and reads inputs
/// [lo_in, hi_in), i.e. a halo of at least ntaps - 1 samples before its range, rounded so that every overlap-save block
/// of the shard coincides with a block of the unsharded stream: the concatenation of the shards equals the single-GPU
/// output bit for bit (tests/test_gpu_sharding.py).
class ShardedFir {
 public:
  size_t lo, hi, lo_in, hi_in, halo;
  ShardedFir(const std::vector<cf32>& taps, size_t total, int rank, int world, int mode = AE_FIR_AUTO) : fir_(taps, 0, mode) {
    if (world <= 0 || rank < 0 || rank >= world) throw Panic(AE_EARG, "bad rank/world");
    lo = total * (size_t)rank / (size_t)world;            // the same split as frame sharding: r n / R .. (r + 1) n / R
    hi = total * (size_t)(rank + 1) / (size_t)world;
    const size_t t1 = taps.size() ? taps.size() - 1 : 0;
    const size_t hop = std::max<size_t>(1, fir_.block_hop());
    // the discarded head must cover the filter memory and, for overlap-save, one whole block (the shard's first block
    // starts from zero history); both ends go outwards to multiples of the hop so that every block of the shard IS a
    // block of the unsharded stream
    const size_t need = std::max(t1, hop > 1 ? hop : (size_t)0);
    lo_in = (lo > need ? lo - need : 0) / hop * hop;
    halo = lo - lo_in;
    hi_in = std::min(total, (hi + hop - 1) / hop * hop);
  }
  /// `in` holds the stream's samples [lo_in, hi_in); returns a view of the hi - lo outputs this rank owns
  DeviceVec filter(const DeviceVec& in, DeviceVec& scratch_out) {
    if (in.len() != hi_in - lo_in) throw Panic(AE_ELEN, "ShardedFir: input must hold the samples [lo_in, hi_in)");
    fir_.reset();
    fir_.filter(in, scratch_out);
    return scratch_out.slice(halo, halo + (hi - lo));
  }

 private:
  Fir fir_;
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

/// pool::{make, Pool, Elem} (src/pool.rs:43-250): reusable, expensive-to-make objects behind guards — here pinned host
/// buffers and device vectors.  take() is the bounded use (nothing when the pool is empty), take_or_make() grows it; the
/// guard resets its element and returns it to the pool when it goes out of scope.
namespace pool {
template <class T>
class Pool;
namespace detail {
template <class T>
struct Inner {
  std::mutex mu;
  std::vector<T> elems;
  std::function<T()> maker;
  std::function<void(T&)> resetter;
  size_t cap = 0;
};
}  // namespace detail
template <class T>
class Elem {
 public:
  Elem(std::shared_ptr<detail::Inner<T>> pool, T val) : pool_(std::move(pool)), val_(std::move(val)), live_(true) {}
  Elem(const Elem&) = delete;
  Elem(Elem&& o) noexcept : pool_(std::move(o.pool_)), val_(std::move(o.val_)), live_(std::exchange(o.live_, false)) {}
  ~Elem() {
    if (!live_) return;
    pool_->resetter(val_);
    std::lock_guard<std::mutex> lk(pool_->mu);
    pool_->elems.push_back(std::move(val_));
  }
  T& operator*() { return val_; }
  T* operator->() { return &val_; }

 private:
  std::shared_ptr<detail::Inner<T>> pool_;
  T val_;
  bool live_;
};
template <class T>
class Pool {
 public:
  explicit Pool(std::shared_ptr<detail::Inner<T>> inner) : inner_(std::move(inner)) {}
  Pool clone() const { return Pool(inner_); }
  std::optional<Elem<T>> take() {
    std::lock_guard<std::mutex> lk(inner_->mu);
    if (inner_->elems.empty()) return std::nullopt;
    T v = std::move(inner_->elems.back());
    inner_->elems.pop_back();
    return Elem<T>(inner_, std::move(v));
  }
  Elem<T> take_or_make() {
    {
      std::lock_guard<std::mutex> lk(inner_->mu);
      if (!inner_->elems.empty()) {
        T v = std::move(inner_->elems.back());
        inner_->elems.pop_back();
        return Elem<T>(inner_, std::move(v));
      }
      inner_->cap += 1;
    }
    return Elem<T>(inner_, inner_->maker());
  }
  size_t len() const { std::lock_guard<std::mutex> lk(inner_->mu); return inner_->elems.size(); }
  bool is_empty() const { return len() == 0; }
  size_t cap() const { std::lock_guard<std::mutex> lk(inner_->mu); return inner_->cap; }

 private:
  std::shared_ptr<detail::Inner<T>> inner_;
};
template <class T>
Pool<T> make(size_t initial_len, std::function<T()> maker, std::function<void(T&)> resetter) {
  auto inner = std::make_shared<detail::Inner<T>>();
  inner->maker = std::move(maker);
  inner->resetter = std::move(resetter);
  for (size_t i = 0; i < initial_len; ++i) {
    T e = inner->maker();
    inner->resetter(e);
    inner->elems.push_back(std::move(e));
  }
  inner->cap = inner->elems.size();
  return Pool<T>(std::move(inner));
}
}  // namespace pool

/// pipeline::{new, Pipeline::add_stage, finish} (src/pipeline.rs:26-137) over ae_pipeline_*: a CUDA stream per stage instead
/// of a thread, CUDA events instead of channels.  Sender::send(item) runs every stage's closure once, in order, on the
/// calling thread; the closures only queue work (kernels, *_async copies) — while one runs, everything the library launches
/// goes to that stage's stream, so stage k of item i overlaps stage k+1 of item i-1.  Receiver::recv() returns finished
/// items in order; report() is the line the reference's stage threads print.
namespace pipeline {
namespace detail {
struct Holder { virtual ~Holder() = default; };
template <class T>
struct Value : Holder {
  T v;
  explicit Value(T x) : v(std::move(x)) {}
};
using Box = std::unique_ptr<Holder>;
struct State;
struct StageCtx { State* st; size_t index; };
struct State {
  ae_pipeline* h = nullptr;
  std::vector<std::function<Box(Box)>> ops;
  std::vector<std::unique_ptr<StageCtx>> ctx;
  std::exception_ptr error;
  ~State() { if (h) ae_pipeline_destroy(h); }
  static ae_status tramp(void* user, size_t /*slot*/, void* item) {
    auto* sc = static_cast<StageCtx*>(user);
    auto* box = static_cast<Box*>(item);
    try {
      *box = sc->st->ops[sc->index](std::move(*box));
      return AE_OK;
    } catch (...) {                       // an exception must not unwind through the C frames
      sc->st->error = std::current_exception();
      return AE_EARG;
    }
  }
  template <class In, class Out, class F>
  void add(const char* name, F op) {
    ops.push_back([op = std::move(op)](Box b) mutable -> Box {
      auto* in = static_cast<Value<In>*>(b.get());
      return Box(new Value<Out>(op(std::move(in->v))));
    });
    ctx.push_back(std::make_unique<StageCtx>(StageCtx{this, ops.size() - 1}));
    check(ae_pipeline_add_stage(h, name, &State::tramp, ctx.back().get()));
  }
};
}  // namespace detail

template <class I>
class Sender {
 public:
  explicit Sender(std::shared_ptr<detail::State> s) : s_(std::move(s)) {}
  void send(I item) {
    auto* box = new detail::Box(new detail::Value<I>(std::move(item)));
    const ae_status st = ae_pipeline_send(s_->h, box);
    if (st != AE_OK) {
      delete box;
      if (s_->error) std::rethrow_exception(std::exchange(s_->error, nullptr));
      check(st);
    }
  }

 private:
  std::shared_ptr<detail::State> s_;
};
template <class O>
class Receiver {
 public:
  explicit Receiver(std::shared_ptr<detail::State> s) : s_(std::move(s)) {}
  O recv() {
    void* p = nullptr;
    check(ae_pipeline_recv(s_->h, &p));
    std::unique_ptr<detail::Box> box(static_cast<detail::Box*>(p));
    return std::move(static_cast<detail::Value<O>*>(box->get())->v);
  }
  size_t in_flight() const { return ae_pipeline_in_flight(s_->h); }
  std::vector<ae_pipe_stage> report(bool reset = false) {
    std::vector<ae_pipe_stage> st(ae_pipeline_stages(s_->h));
    check(ae_pipeline_report(s_->h, st.data(), st.size(), reset));
    return st;
  }

 private:
  std::shared_ptr<detail::State> s_;
};
template <class I, class O>
class Pipeline {
 public:
  explicit Pipeline(std::shared_ptr<detail::State> s) : s_(std::move(s)) {}
  /// add another stage: op maps what the previous stage returned to what the next one gets (FnMut(O) -> U)
  template <class F, class U = std::invoke_result_t<F, O>>
  Pipeline<I, U> add_stage(const char* name, F op) {
    s_->template add<O, U>(name, std::move(op));
    return Pipeline<I, U>(std::move(s_));
  }
  std::pair<Sender<I>, Receiver<O>> finish() { return {Sender<I>(s_), Receiver<O>(s_)}; }

 private:
  std::shared_ptr<detail::State> s_;
};
/// pipeline::new(name, op): `depth` items may be in flight (the reference's channels are unbounded; here the device
/// buffers an item carries bound it — send() waits for the oldest item when `depth` are in flight)
template <class I, class F, class O = std::invoke_result_t<F, I>>
Pipeline<I, O> new_(const char* name, F op, int depth = 3) {
  auto s = std::make_shared<detail::State>();
  check(ae_pipeline_create(depth, &s->h));
  s->template add<I, O>(name, std::move(op));
  return Pipeline<I, O>(std::move(s));
}
}  // namespace pipeline

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
