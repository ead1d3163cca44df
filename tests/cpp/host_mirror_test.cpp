// C++ host-mirror test: the reference crate's own unit tests and doctests for the hot path,
// re-stated against aether_b200.hpp (each block cites the Rust test it mirrors).  Exit code 0 = pass.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "aether_b200.hpp"

using namespace aether;

// assert_evm! (src/lib.rs:26-49), literal formula
static void assert_evm(const std::vector<cf32>& act, const std::vector<cf32>& ref, double db, const char* what) {
  if (act.size() != ref.size()) { std::printf("FAIL %s: Input slices/vectors must be same length\n", what); std::exit(1); }
  const float thr = (float)std::pow(10.0, db / 10.0);
  for (size_t i = 0; i < act.size(); ++i) {
    const float evm = std::abs(act[i] - ref[i]);
    const float limit = std::abs(ref[i]) * thr;
    if (evm > limit) { std::printf("FAIL %s: EVM limit exceeded at element %zu: %g > %g\n", what, i, evm, limit); std::exit(1); }
  }
}
static std::vector<cf32> rep(cf32 v, size_t n) { return std::vector<cf32>(n, v); }
#define EXPECT_PANIC(stmt, st, what)                                                     \
  do {                                                                                   \
    bool ok__ = false;                                                                   \
    try { stmt; } catch (const Panic& p) { ok__ = (p.status == (st)); }                  \
    if (!ok__) { std::printf("FAIL %s: expected panic\n", what); std::exit(1); }         \
  } while (0)

int main() {
  init(0);
  {  // src/vecops.rs:340-424
    DeviceVec v(rep({0.5f, 0.5f}, 100));
    assert_evm(v.vec_scale(2.0f).to_host(), rep({1, 1}, 100), -80, "vec_scale");
    DeviceVec ones(rep({1, 1}, 100)), twos(rep({0, 2}, 100));
    assert_evm(ones.vec_mul(twos).to_host(), rep({-2, 2}, 100), -80, "vec_mul");
    DeviceVec a(rep({2, 2}, 100)), b(rep({2, 0}, 100));
    assert_evm(a.vec_div(b).to_host(), rep({1, 1}, 100), -80, "vec_div");
    DeviceVec c(rep({1, 1}, 100));
    assert_evm(c.vec_conj().to_host(), rep({1, -1}, 100), -80, "vec_conj");
    DeviceVec even(std::vector<cf32>{{0, 0}, {1, 0}, {2, 0}, {3, 0}});
    assert_evm(even.vec_mirror().to_host(), {{2, 0}, {3, 0}, {0, 0}, {1, 0}}, -80, "vec_mirror");
    DeviceVec short_v(4);
    EXPECT_PANIC(c.vec_add(short_v), AE_ELEN, "Vectors must have same length");
  }
  {  // VecOps doctest src/vecops.rs:19-36
    DeviceVec v(rep({2, 2}, 100)), twos(rep({2, 2}, 100)), ones(rep({1, 1}, 100));
    v.vec_div(twos).vec_mul(twos).vec_zero().vec_add(ones).vec_sub(twos).vec_clone(ones)
        .vec_mutate([](cf32& z) { z.imag(-1.0f); }).vec_conj().vec_mirror();
    assert_evm(v.to_host(), rep({1, 1}, 100), -80, "vecops doctest chain");
  }
  {  // vec_mutate src/vecops.rs:427-441
    DeviceVec v(rep({1, 1}, 100));
    int x = 0;
    v.vec_mutate([&x](cf32& z) { z *= (float)x; ++x; });
    std::vector<cf32> lin(100);
    for (int i = 0; i < 100; ++i) lin[i] = {(float)i, (float)i};
    assert_evm(v.to_host(), lin, -80, "vec_mutate");
  }
  {  // fft doctest src/fft.rs:93-117
    DeviceVec data(rep({1, 0}, 128));
    data.vec_fft(Scale::None());
    std::vector<cf32> right(128, {0, 0});
    right[0] = {128, 0};
    assert_evm(data.to_host(), right, -80, "fft doctest: DC bin");
    Cfft f = Cfft::with_len(128);
    f.ibwd(data, Scale::N());
    assert_evm(data.to_host(), rep({1, 0}, 128), -80, "fft doctest: ibwd N");
    data.vec_rfft(f, Scale::SN()).vec_scale(2.0f).vec_rifft(f, Scale::SN());
    assert_evm(data.to_host(), rep({2, 0}, 128), -72, "fft doctest: rfft*2*rifft");
    DeviceVec wrong(100);
    EXPECT_PANIC(f.ifwd(wrong, Scale::None()), AE_ELEN, "Input and FFT must be the same length");
  }
  {  // src/modulation.rs:158-196
    Modulation q = qpsk(), b = bpsk();
    DeviceBits bits(std::vector<uint8_t>{0, 0, 1, 0, 0, 1, 1, 1});
    assert_evm(q.modulate(bits).to_host(), {{1, 1}, {-1, 1}, {1, -1}, {-1, -1}}, -80, "generic_qpsk");
    DeviceBits bb(std::vector<uint8_t>{0, 1, 0, 1});
    assert_evm(b.modulate(bb).to_host(), {{1, 1}, {-1, -1}, {1, 1}, {-1, -1}}, -80, "generic_bpsk");
    DeviceBits zeros(std::vector<uint8_t>(100, 0)), out(100);
    DeviceVec sym = q.modulate(zeros);
    q.demod_naive(sym, out);
    if (out.to_host() != std::vector<uint8_t>(100, 0)) { std::printf("FAIL naive_demod\n"); return 1; }
  }
  {  // src/sampling.rs:73-169, src/sequence.rs:61-68
    DeviceVec src(std::vector<cf32>{{0, 0}, {3, 3}, {6, 6}, {9, 9}});
    DeviceVec dst = DeviceVec::with_capacity(16);
    sampling::interpolate(src, dst, 2);
    std::vector<cf32> chk(10);
    for (int i = 0; i < 10; ++i) chk[i] = {(float)i, (float)i};
    if (dst.to_host() != chk) { std::printf("FAIL interpolate_2_between\n"); return 1; }
    DeviceVec s7(7), d3(3);
    EXPECT_PANIC(sampling::downsample(s7, d3), AE_ELEN, "downsample_7_v_3_fail");
    if (sequence::generate({1, 0}, {1, 2}, 6).to_host() != std::vector<uint8_t>{1, 0, 1, 1, 0, 1}) { std::printf("FAIL simple_sequence\n"); return 1; }
  }
  {  // examples/modem.rs:15-32
    Modulation m = qpsk();
    std::vector<uint8_t> b(100);
    for (int i = 0; i < 100; ++i) b[i] = (uint8_t)((i * 7 + i / 3) & 1);
    DeviceBits bits(b);
    DeviceVec sym = m.modulate(bits);
    Awgn n = noise::new_(0.01f, 815);
    n.apply(sym);
    DeviceBits rx(100);
    m.demod_naive(sym, rx, Compat::Corrected);
    if (rx.to_host() != b) { std::printf("FAIL modem loop-back\n"); return 1; }
  }
  {  // headline chain: device call, streamed host blocks (pipeline.rs / pool.rs analogue) and VecStats agree
    const size_t n = 1024, block = 8, blocks = 5;
    std::vector<cf32> taps(64), x(n * block * blocks);
    for (size_t k = 0; k < taps.size(); ++k) taps[k] = cf32(1.f / (1.f + k), 0.01f * k);
    uint32_t lcg = 12345;
    for (auto& e : x) {
      lcg = lcg * 1664525u + 1013904223u; const float re = (float)(lcg >> 8) / 8388608.f - 1.f;
      lcg = lcg * 1664525u + 1013904223u; const float im = (float)(lcg >> 8) / 8388608.f - 1.f;
      e = cf32(re, im);
    }
    FftFirDemod chain(n, taps, Scale::SN());
    DeviceVec dx(x);
    DeviceBits want(2 * x.size());
    chain.run(dx, want);
    const std::vector<uint8_t> w = want.to_host();
    ChainPipeline pipe(chain, block, 2);
    std::vector<uint8_t> got(2 * x.size());
    for (size_t b = 0; b < blocks; ++b) pipe.send(x.data() + b * block * n, got.data() + 2 * b * block * n);
    for (size_t b = 0; b < blocks; ++b)
      if (pipe.recv() != got.data() + 2 * b * block * n) { std::printf("FAIL pipeline order\n"); return 1; }
    if (got != w) { std::printf("FAIL pipeline bits\n"); return 1; }
    const auto rep = pipe.report();
    if (rep[1].processed != blocks || rep[1].utilisation_pct <= 0) { std::printf("FAIL pipeline report\n"); return 1; }
    const ae_vecstats st = dx.vec_stats();
    size_t imax = 0;
    float vmax = -1.f;
    for (size_t i = 0; i < x.size(); ++i) {
      const float a = x[i].real() * x[i].real(), b2 = x[i].imag() * x[i].imag(), v = a + b2;
      if (v > vmax) { vmax = v; imax = i; }
    }
    if (st.n != x.size() || st.max_idx != imax || st.max_val != vmax) { std::printf("FAIL vec_stats\n"); return 1; }
  }
  {  // multi-GPU and fused surface: counters, communicator (one rank), graph replay, sharded FIR, modem / OFDM kernels
    Modulation m = qpsk();
    Awgn g = noise::generator();
    g.set_power(0.5f);
    DeviceBits tx = sequence::generate(sequence::expand(1, 31).to_host(), {28, 31}, 4096), rx(4096);   // src/sequence.rs:41-46
    DeviceStats st;
    chain::modem_fused(m, g, tx, rx, &st, Compat::Corrected);
    if (g.tell() != 2048) { std::printf("FAIL awgn tell after modem_fused\n"); return 1; }
    ae_stats a = st.read();
    DeviceStats st2;
    st2.count_bit_errors(tx, rx);
    if (a.n_bits != 4096 || a.bit_errors == 0 || a.bit_errors != st2.read().bit_errors) { std::printf("FAIL modem_fused counters\n"); return 1; }
    // the same noise stream again: seek back, apply on the stand-alone kernels
    g.seek(0);
    DeviceVec sym = m.modulate(tx);
    g.apply(sym, Compat::Corrected);             // compat=reference scales the noise twice (src/noise.rs:58)
    DeviceBits rx2(4096);
    m.demod_naive(sym, rx2, Compat::Corrected);
    if (rx2.to_host() != rx.to_host()) { std::printf("FAIL modem_fused vs stand-alone kernels\n"); return 1; }
    // OFDM counters: two half-runs with global frame ids == one run; all-reduce over a one-rank communicator is the identity
    DeviceStats whole, halves;
    chain::ofdm_chain(2048, 8, 0, 0.5f, 7, &whole);
    chain::ofdm_chain(2048, 4, 0, 0.5f, 7, &halves);
    chain::ofdm_chain(2048, 4, 4, 0.5f, 7, &halves);
    const ae_stats w = whole.read(), h = halves.read();
    if (w.bit_errors != h.bit_errors || w.n_bits != h.n_bits || w.n_bits != 2u * 2048u * 8u) { std::printf("FAIL ofdm_chain counters\n"); return 1; }
    Comm c = Comm::init_rank(Comm::unique_id(), 1, 0);
    if (c.info().nranks != 1 || c.info().rank != 0) { std::printf("FAIL comm info\n"); return 1; }
    c.allreduce(whole);
    if (whole.read().bit_errors != w.bit_errors) { std::printf("FAIL one-rank allreduce\n"); return 1; }
    // CUDA graph: three recorded vec_scale launches, replayed twice
    DeviceVec v(rep({1, 1}, 256));
    Graph gr = Graph::record([&] { for (int i = 0; i < 3; ++i) v.vec_scale(2.0f).flush(); });
    gr.launch();
    gr.launch();
    assert_evm(v.to_host(), rep({64, 64}, 256), -80, "graph replay");
    // sharded FIR: three shards of one stream == the unsharded filter, bit for bit
    std::vector<cf32> taps(64), x(40000);
    for (size_t k = 0; k < taps.size(); ++k) taps[k] = cf32(1.f / (1.f + k), 0.01f * k);
    uint32_t lcg = 99;
    for (auto& e : x) { lcg = lcg * 1664525u + 1013904223u; e = cf32((float)(lcg >> 8) / 8388608.f - 1.f, (float)(lcg & 0xffff) / 32768.f - 1.f); }
    Fir whole_f(taps, 0, AE_FIR_OVERLAP_SAVE);
    DeviceVec dx(x), dy(x.size());
    whole_f.filter(dx, dy);
    const std::vector<cf32> want = dy.to_host();
    for (int r = 0; r < 3; ++r) {
      ShardedFir sf(taps, x.size(), r, 3, AE_FIR_OVERLAP_SAVE);
      DeviceVec in(std::vector<cf32>(x.begin() + sf.lo_in, x.begin() + sf.hi_in)), work(sf.hi_in - sf.lo_in);
      const std::vector<cf32> got = sf.filter(in, work).to_host();
      if (got.size() != sf.hi - sf.lo || std::memcmp(got.data(), want.data() + sf.lo, got.size() * sizeof(cf32)) != 0) {
        std::printf("FAIL sharded FIR, shard %d\n", r);
        return 1;
      }
    }
    // spectrogram of a DC frame: all power in the middle bin; correlator of a frame with itself peaks at lag 0
    Cfft f = Cfft::with_len(1024);
    DeviceVec dc(rep({1, 0}, 1024));
    DeviceF32 lev(1024);
    f.spectrogram(dc, lev, false);
    const std::vector<float> l = lev.to_host();
    size_t arg = 0;
    for (size_t i = 0; i < l.size(); ++i) if (l[i] > l[arg]) arg = i;
    if (arg != 512) { std::printf("FAIL spectrogram DC bin at %zu\n", arg); return 1; }
    // benches/benches.rs:410-416: input.vec_rfft(SN).vec_mul(&sig).vec_rifft(SN); a flat spectrum of 2 doubles the frame
    DeviceVec p(std::vector<cf32>(x.begin(), x.begin() + 1024)), flat(rep({2, 0}, 1024));
    f.correlate(p, flat, Scale::SN());
    std::vector<cf32> twice(x.begin(), x.begin() + 1024);
    for (auto& e : twice) e *= 2.0f;
    {
      double err = 0, ref = 0;
      const std::vector<cf32> got = p.to_host();
      for (size_t i = 0; i < got.size(); ++i) { err += std::norm(got[i] - twice[i]); ref += std::norm(twice[i]); }
      if (!(err <= 1e-10 * ref)) { std::printf("FAIL correlator round trip: %g\n", err / ref); return 1; }
    }
    if (version().empty() || device_count() < 1 || sm_count() < 1 || launch_count() == 0) { std::printf("FAIL runtime info\n"); return 1; }
  }
  {  // src/pool.rs:264-329: taking, resetting, taking_or_making
    auto p = pool::make<std::vector<uint8_t>>(1, [] { std::vector<uint8_t> v; v.reserve(50); return v; }, [](std::vector<uint8_t>& o) { o.clear(); });
    if (p.len() != 1 || p.cap() != 1) { std::printf("FAIL pool make\n"); return 1; }
    {
      auto c1 = p.take();
      if (!c1 || p.len() != 0 || p.cap() != 1) { std::printf("FAIL pool: first checkout\n"); return 1; }
      for (int x = 0; x < 50; ++x) (**c1).push_back((uint8_t)x);
    }
    if (p.len() != 1 || p.cap() != 1) { std::printf("FAIL pool: give back\n"); return 1; }
    {
      auto c1 = p.take();
      auto c2 = p.take();
      if (!c1 || c2 || !(**c1).empty()) { std::printf("FAIL pool: second / third checkout, resetter\n"); return 1; }
    }
    auto q = pool::make<std::vector<uint8_t>>(0, [] { return std::vector<uint8_t>(); }, [](std::vector<uint8_t>& o) { o.clear(); });
    {
      auto e1 = q.take_or_make();
      if (q.len() != 0 || q.cap() != 1) { std::printf("FAIL pool: take_or_make 1\n"); return 1; }
      auto e2 = q.clone().take_or_make();
      if (q.len() != 0 || q.cap() != 2) { std::printf("FAIL pool: take_or_make 2\n"); return 1; }
    }
    if (q.len() != 2 || q.cap() != 2 || q.is_empty()) { std::printf("FAIL pool: both guards dropped\n"); return 1; }
  }
  {  // src/pipeline.rs: three stages (H2D copy, transform + VecOps + demod, D2H copy) on three streams, pooled pinned and
     // device buffers travelling through them; results == the same calls made one by one, in order
    struct Block {
      PinnedBuf<cf32> h_in;
      PinnedBuf<uint8_t> h_bits;
      DeviceVec d;
      DeviceBits bits;
      int tag = -1;
      explicit Block(size_t n) : h_in(n), h_bits(2 * n), d(n), bits(2 * n) {}
    };
    const size_t n = 1024, frames = 16, ns = n * frames, blocks = 6;
    const int depth = 3;
    auto blocks_pool = pool::make<Block>(0, [&] { return Block(ns); }, [](Block& b) { b.tag = -1; b.bits.clear(); });   // demod_naive appends
    Cfft fft = Cfft::with_len(n);
    Modulation q = qpsk();
    DeviceVec w(rep({0.6f, 0.8f}, ns));
    using Item = pool::Elem<Block>;
    auto built = pipeline::new_<Item>("h2d", [&](Item e) { e->d.upload_async(e->h_in.data(), ns); return e; }, depth)
                     .add_stage("fft-mul-demod", [&](Item e) {
                       fft.ifwd(e->d, Scale::SN(), frames);
                       e->d.vec_mul(w).vec_conj();
                       q.demod_naive(e->d, e->bits);
                       return e;
                     })
                     .add_stage("d2h", [&](Item e) { e->bits.download_async(e->h_bits.data(), 2 * ns); return e; })
                     .finish();
    auto& tx = built.first;
    auto& rx = built.second;
    std::vector<std::vector<cf32>> xs(blocks, std::vector<cf32>(ns));
    uint32_t lcg = 4242;
    for (auto& x : xs) for (auto& e : x) { lcg = lcg * 1664525u + 1013904223u; e = cf32((float)(lcg >> 8) / 8388608.f - 1.f, (float)(lcg & 0xffff) / 32768.f - 1.f); }
    std::vector<std::pair<int, std::vector<uint8_t>>> got;
    auto take_result = [&] {
      Item e = rx.recv();
      got.emplace_back(e->tag, std::vector<uint8_t>(e->h_bits.data(), e->h_bits.data() + 2 * ns));
    };
    for (size_t i = 0; i < blocks; ++i) {
      if (rx.in_flight() == (size_t)depth) take_result();
      Item e = blocks_pool.take_or_make();
      e->tag = (int)i;
      std::memcpy(e->h_in.data(), xs[i].data(), ns * sizeof(cf32));
      tx.send(std::move(e));
    }
    while (rx.in_flight()) take_result();
    if (blocks_pool.cap() > (size_t)depth || blocks_pool.len() != blocks_pool.cap()) { std::printf("FAIL pipeline: pool reuse\n"); return 1; }
    for (size_t i = 0; i < blocks; ++i) {
      if (got[i].first != (int)i) { std::printf("FAIL pipeline: order\n"); return 1; }
      DeviceVec d(xs[i]);
      fft.ifwd(d, Scale::SN(), frames);
      d.vec_mul(w).vec_conj();
      DeviceBits bits(2 * ns);
      q.demod_naive(d, bits);
      if (bits.to_host() != got[i].second) { std::printf("FAIL pipeline: block %zu differs from the serial calls\n", i); return 1; }
    }
    const auto rep3 = rx.report();
    if (rep3.size() != 3 || std::string(rep3[1].name) != "fft-mul-demod" || rep3[0].processed != blocks || !(rep3[2].active_ms > 0)) {
      std::printf("FAIL pipeline: report\n");
      return 1;
    }
    // a throwing stage surfaces from send() and leaves nothing in flight
    auto bad = pipeline::new_<int>("ok", [](int v) { return v + 1; }, 2).add_stage("boom", [](int) -> int { throw std::runtime_error("stage failed"); }).finish();
    bool threw = false;
    try { bad.first.send(1); } catch (const std::runtime_error&) { threw = true; }
    if (!threw || bad.second.in_flight() != 0) { std::printf("FAIL pipeline: stage error\n"); return 1; }
  }
  sync();
  std::printf("host mirror ok\n");
  return 0;
}
