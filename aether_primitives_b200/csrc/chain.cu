// chain.cu — fused chains (sm_100a):
//   K14 fft_fir_demod : per frame Cfft::fwd(scale) -> FIR (zero state per frame) -> QPSK demod_naive
//                       8 B in + 2 B out per sample; nothing else touches HBM.
//   K12 ofdm_chain    : M-sequence -> QPSK -> bwd FFT(SN) -> AWGN -> fwd FFT(SN) -> demod -> BER/EVM
//
// K14 algorithm.  A direct 64-tap FIR costs 512 flop/sample and would make the
// chain FP32-bound at ~1/4 of the HBM roofline.  Instead use the convolution theorem on the
// transform itself.  With X = DFT_s(x) (exponent sign s), the CIRCULAR convolution of X with h is
//     C[n] = sum_k h[k] X[(n-k) mod N] = DFT_s( x .* w )[n],   w[m] = sum_k h[k] exp(-s 2 pi i m k/N),
// and the linear, zero-state FIR the reference semantics ask for differs from it only in the first
// T-1 outputs:
//     y[n] = C[n] - sum_{k=n+1}^{T-1} h[k] X[N + n - k]        (n < T-1),
// which needs the last T-1 bins of the plain transform.  So each frame runs two in-register FFTs
// (A = DFT(x), B = DFT(x .* w)), a triangular fix-up of T(T-1)/2 complex MACs, and the exact
// reference hard decision on the result:  ~2*50 + 6 + 16 flop/sample instead of 562.
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "async_copy.cuh"
#include "fft_device.cuh"
#include "internal.h"

namespace ae {

// frame slots per CTA: 128-thread CTAs (256 when one frame needs them), 4 (2) CTAs per SM
template <int N>
struct ChainLaunch {
  static constexpr int T = FftCfg<N>::T;
  static constexpr int F = T >= 128 ? 1 : (128 / T);
  static constexpr int THREADS = F * T;
  static constexpr int MINB = THREADS <= 128 ? 4 : 2;
};
// shared-memory layout (in cf32): [taps zero padded: HP][taps shifted by one: HP][per slot: A | B | rt: RP | fixA: RP | fixB: RP | staging: N if STAGED][mbarriers]
__host__ __device__ constexpr int chain_hp(int ntaps) { return ((ntaps + 68 + 1) / 2) * 2; }
__host__ __device__ constexpr int chain_rp(int ntaps) { return ((ntaps + 4 + 1) / 2) * 2; }
template <int N>
__host__ __device__ constexpr size_t chain_slot_elems(int ntaps, bool staged) {
  return 2 * (size_t)FftCfg<N>::SMEM_ELEMS + 3 * (size_t)chain_rp(ntaps) + (staged ? N : 0);
}

// PRUNE : ntaps-1 <= N/16, so the fix-up only needs register 15 (positions >= N - N/16) of A and
//         transform A is pruned to those bins (fft_device.cuh, TAIL0).
// STAGED: the next frame's 8N input bytes are fetched by ONE cp.async.bulk (TMA) per frame into the
//         slot's staging buffer while the current frame is transformed; completion arrives on an
//         mbarrier.  Needs a 16-byte aligned input; otherwise the plain-load variant runs.
template <int N, bool INV, bool PRUNE, bool STAGED>
__global__ void __launch_bounds__(ChainLaunch<N>::THREADS, ChainLaunch<N>::MINB)
chain_fused_kernel(const float2* __restrict__ x, uint8_t* __restrict__ bits, size_t frames, const float2* __restrict__ window,
                   const float2* __restrict__ taps, int ntaps, const float2* __restrict__ tw, float scale, int compat) {
  using C = FftCfg<N>;
  using LC = ChainLaunch<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* hs = reinterpret_cast<float2*>(smem_raw);
  const int f = threadIdx.x / C::T;
  const int t = threadIdx.x % C::T;
  const int HP = chain_hp(ntaps), RP = chain_rp(ntaps);
  const size_t slot = chain_slot_elems<N>(ntaps, STAGED);
  float2* hso = hs + HP;              // hso[j] = h[j+1]: makes the pair (h[j+1], h[j+2]) 16-byte aligned for even j
  float2* smA = hs + 2 * HP + (size_t)f * slot;
  float2* smB = smA + C::SMEM_ELEMS;
  float2* rt = smB + C::SMEM_ELEMS;   // rt[i] = scale * A[N-1-i], zero padded
  float2* fix = rt + RP;              // fix-up partial sums A (and the whole sum on the generic path)
  float2* fixb = fix + RP;            // fix-up partial sums B
  float2* xin = fixb + RP;            // STAGED only
  uint64_t* bar = reinterpret_cast<uint64_t*>(hs + 2 * HP + (size_t)LC::F * slot) + f;
  for (int i = threadIdx.x; i < HP; i += LC::THREADS) {
    hs[i] = i < ntaps ? __ldg(taps + i) : make_float2(0.0f, 0.0f);
    hso[i] = i + 1 < ntaps ? __ldg(taps + i + 1) : make_float2(0.0f, 0.0f);
  }
  for (int i = t; i < RP; i += C::T) rt[i] = make_float2(0.0f, 0.0f);
  const size_t stride = (size_t)gridDim.x * LC::F;
  size_t frame = (size_t)blockIdx.x * LC::F + f;
  if (STAGED && t == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    if (frame < frames) {
      mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
      bulk_g2s(xin, x + frame * N, N * (uint32_t)sizeof(float2), bar);
    }
  }
  __syncthreads();
  const int tm1 = ntaps - 1;
  const unsigned hi_shift = compat == AE_COMPAT_REFERENCE ? 9u : 8u;  // second byte: idx & 2 vs (idx >> 1) & 1
  uint32_t phase = 0;
  // persistent: every frame slot (C::T threads) walks its own frames and only ever synchronises
  // with itself, so the slots of an SM drift out of phase and overlap their load / compute phases
  for (; frame < frames; frame += stride) {
    // ab[0] = x (-> A = DFT(x)), ab[1] = x .* w (-> B = circular convolution of scale*X with h);
    // both transforms advance pass by pass together: twiddles loaded once, barriers shared
    float2 ab[2][16];
    if (STAGED) {
      mbar_wait(bar, phase);
      phase ^= 1u;
#pragma unroll
      for (int m = 0; m < 16; ++m) ab[0][m] = xin[t + m * C::T];
    } else {
      const float2* src = x + frame * N;
#pragma unroll
      for (int m = 0; m < 16; ++m) ab[0][m] = ld_stream(src + t + m * C::T);
    }
#pragma unroll
    for (int m = 0; m < 16; ++m) ab[1][m] = cx_mul(ab[0][m], __ldg(window + t + m * C::T));
    float2* const sm2[2] = {smA, smB};
    // after the first slot barrier every thread has consumed its staged inputs: refill the buffer
    auto prefetch = [&]() {
      if (STAGED && t == 0 && frame + stride < frames) {
        mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
        bulk_g2s(xin, x + (frame + stride) * N, N * (uint32_t)sizeof(float2), bar);
      }
    };
    fft_frames<N, INV, 2, PRUNE>(ab, sm2, tw, t, f, prefetch);

    // only the last T-1 bins of A are needed (scaled like Cfft::fwd's output); stored REVERSED so the
    // fix-up reads rt[i] as a broadcast and the taps contiguously
#pragma unroll
    for (int m = PRUNE ? 15 : 0; m < 16; ++m) {
      const int pos = t + m * C::T;
      if (pos >= N - tm1) rt[N - 1 - pos] = cx_scale_exact(ab[0][m], scale);
    }
    frame_sync<C::T>(f);
    // wrap-around terms of the circular convolution, outputs n < T-1:
    //   fix[n] = sum_{k=n+1}^{T-1} h[k] A[N+n-k] = sum_{i < T-1-n} h[n+1+i] rt[i]
    // Taps and rt are zero padded, so every lane of a warp runs the same trip count with no predicates.
    uint16_t* out = reinterpret_cast<uint16_t*>(bits + 2 * frame * (size_t)N);
    unsigned risky = 0;
    if (PRUNE) {
      // (ntaps-1 <= C::T)  The fix-up phase is bound by shared-memory wavefronts, so each thread owns TWO
      // consecutive outputs (2u, 2u+1): their tap windows overlap, one 16-byte tap read and one 16-byte
      // broadcast rt read feed four complex MACs (half the wavefronts per MAC of one output per lane).
      // The first half of the slot's threads sums taps i < H, the second half taps i >= H.
      const int half = C::T / 2;
      const int u = t < half ? t : t - half;
      const int H = ((tm1 / 2) + 1) & ~1;                       // even split point
      const int lane0_u = u - (int)(threadIdx.x & 31 & (half - 1 < 31 ? half - 1 : 31));
      const int n_first = 2 * (lane0_u > 0 ? lane0_u : 0);       // first output handled by this warp
      const int maxlen = tm1 - n_first;                          // longest sum of the warp
      const int i0 = t < half ? 0 : H;
      const int i1 = t < half ? (H < maxlen ? H : maxlen) : maxlen;
      if (2 * u < tm1) {
        float2 w0a = make_float2(0.0f, 0.0f), w0b = w0a, w1a = w0a, w1b = w0a;
        const float2* hp = hso + 2 * u;                          // hp[j] = h[2u+1+j]
        float4 g = *reinterpret_cast<const float4*>(hp + i0);    // (h[n0+1+i], h[n0+2+i])
        for (int i = i0; i < i1; i += 2) {
          const float4 gn = *reinterpret_cast<const float4*>(hp + i + 2);
          const float4 r = *reinterpret_cast<const float4*>(rt + i);
          const float2 ga = make_float2(g.x, g.y), gb = make_float2(g.z, g.w), gc = make_float2(gn.x, gn.y);
          const float2 r0 = make_float2(r.x, r.y), r1 = make_float2(r.z, r.w);
          cx_fma(w0a, ga, r0);
          cx_fma(w1a, gb, r0);
          cx_fma(w0b, gb, r1);
          cx_fma(w1b, gc, r1);
          g = gn;
        }
        float2* dstp = (t < half ? fix : fixb) + 2 * u;
        *reinterpret_cast<float4*>(dstp) = make_float4(w0a.x + w0b.x, w0a.y + w0b.y, w1a.x + w1b.x, w1a.y + w1b.y);
      }
      // meanwhile-independent part: outputs of registers 1..15 never need the fix-up
#pragma unroll
      for (int m = 1; m < 16; ++m) {
        if (!qpsk_fast_ok(ab[1][m])) risky |= 1u << m;
        out[t + m * C::T] = (uint16_t)qpsk_pair_from_signs(ab[1][m], hi_shift);
      }
      frame_sync<C::T>(f);
      if (t < tm1) {
        const float2 fa = fix[t], fb = fixb[t];
        ab[1][0] = cx_sub(ab[1][0], cx_add(fa, fb));
      }
      if (!qpsk_fast_ok(ab[1][0])) risky |= 1u;
      out[t] = (uint16_t)qpsk_pair_from_signs(ab[1][0], hi_shift);
    } else {
      // generic path (ntaps-1 > C::T): one output per thread and pass, 4 taps per iteration.
      // Output n belongs to thread n mod C::T, which is also the thread that consumes it below.
      for (int n = t; n < tm1; n += C::T) {
        const int n_first = n - (threadIdx.x & 31);  // first output of this warp
        const int maxlen = tm1 - (n_first > 0 ? n_first : 0);
        float2 a0 = make_float2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0;
        const float2* hp = hs + n + 1;
        for (int i = 0; i < maxlen; i += 4) {
          const float4 r01 = *reinterpret_cast<const float4*>(rt + i);
          const float4 r23 = *reinterpret_cast<const float4*>(rt + i + 2);
          cx_fma(a0, hp[i], make_float2(r01.x, r01.y));
          cx_fma(a1, hp[i + 1], make_float2(r01.z, r01.w));
          cx_fma(a2, hp[i + 2], make_float2(r23.x, r23.y));
          cx_fma(a3, hp[i + 3], make_float2(r23.z, r23.w));
        }
        fix[n] = cx_add(cx_add(a0, a1), cx_add(a2, a3));
      }
      // hard decisions (src/modulation.rs:33-56).  Fast tie-free test per symbol; the rare symbols that
      // fail it (near an axis, NaN, inf) are collected in a mask and re-decided exactly afterwards.
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const int n = t + m * C::T;
        if (n < tm1) ab[1][m] = cx_sub(ab[1][m], fix[n]);
        if (!qpsk_fast_ok(ab[1][m])) risky |= 1u << m;
        out[n] = (uint16_t)qpsk_pair_from_signs(ab[1][m], hi_shift);
      }
    }
    if (risky) {
#pragma unroll
      for (int m = 0; m < 16; ++m)
        if (risky & (1u << m)) out[t + m * C::T] = (uint16_t)qpsk_pair_from_index(demod_qpsk_exact_slow(ab[1][m]), hi_shift);
    }
  }  // frame loop
}

// dynamic shared memory of chain_fused_kernel<nfft> (layout above); must stay below the 227 KB opt-in limit
static size_t chain_smem_bytes(size_t nfft, size_t ntaps, bool staged) {
  const size_t T = nfft / 16, F = T >= 128 ? 1 : 128 / T;
  const size_t slot = 2 * (nfft + nfft / 16) + 3 * (size_t)chain_rp((int)ntaps) + (staged ? nfft : 0);
  return (2 * (size_t)chain_hp((int)ntaps) + F * slot) * sizeof(float2) + F * sizeof(uint64_t);
}
constexpr size_t kSmemOptinLimit = 227 * 1024;
bool chain_fused_supported(size_t nfft, size_t ntaps) {
  return nfft >= 256 && nfft <= 4096 && (nfft & (nfft - 1)) == 0 && ntaps >= 1 && ntaps <= nfft &&
         chain_smem_bytes(nfft, ntaps, false) <= kSmemOptinLimit;   // else the plan takes the unfused path
}

template <int N, bool STAGED, class K>
static void launch_chain_kernel(K kern, const float2* x, uint8_t* bits, size_t frames, const float2* window, const float2* taps,
                                size_t ntaps, const float2* tw, float scale, int compat, cudaStream_t st) {
  using LC = ChainLaunch<N>;
  const size_t smem = (2 * (size_t)chain_hp((int)ntaps) + (size_t)LC::F * chain_slot_elems<N>((int)ntaps, STAGED)) * sizeof(float2) +
                      (size_t)LC::F * sizeof(uint64_t);
  const size_t want = (frames + LC::F - 1) / LC::F;
  // grid = SM count x resident CTAs; queried once per (kernel instantiation, shared-memory size)
  static thread_local size_t cached_smem = 0, cached_resident = 0;
  static thread_local const void* cached_kern = nullptr;
  static thread_local int cached_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cached_smem != smem || cached_kern != (const void*)kern || cached_dev != dev) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1, sms = 148;
    cached_dev = dev;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, LC::THREADS, smem);
    cached_resident = (size_t)sms * (per_sm > 0 ? per_sm : 1);
    cached_smem = smem;
    cached_kern = (const void*)kern;
  }
  const unsigned grid = (unsigned)(want < cached_resident ? want : cached_resident);
  kern<<<grid, LC::THREADS, smem, st>>>(x, bits, frames, window, taps, (int)ntaps, tw, scale, compat);
}

template <int N>
static void launch_chain_n(const float2* x, uint8_t* bits, size_t frames, const float2* window, const float2* taps,
                           size_t ntaps, const float2* tw, bool inverse, float scale, int compat, cudaStream_t st) {
  const bool prune = ntaps - 1 <= (size_t)FftCfg<N>::T;
  static const char* no_tma = getenv("AE_CHAIN_NO_TMA");
  const bool staged = ((uintptr_t)x % 16) == 0 && !no_tma && chain_smem_bytes(N, ntaps, true) <= kSmemOptinLimit;
#define AE_GO(I, P, S) launch_chain_kernel<N, S>(chain_fused_kernel<N, I, P, S>, x, bits, frames, window, taps, ntaps, tw, scale, compat, st)
  if (inverse) {
    if (prune) { if (staged) AE_GO(true, true, true); else AE_GO(true, true, false); }
    else { if (staged) AE_GO(true, false, true); else AE_GO(true, false, false); }
  } else {
    if (prune) { if (staged) AE_GO(false, true, true); else AE_GO(false, true, false); }
    else { if (staged) AE_GO(false, false, true); else AE_GO(false, false, false); }
  }
#undef AE_GO
}

void launch_chain_fused(const float2* x, uint8_t* bits, size_t nfft, size_t frames, const float2* window, const float2* taps,
                        size_t ntaps, const float2* tw, bool inverse, float scale, int compat, cudaStream_t st) {
  if (frames == 0) return;
  switch (nfft) {
#define AE_CASE(NN) case NN: launch_chain_n<NN>(x, bits, frames, window, taps, ntaps, tw, inverse, scale, compat, st); break;
    AE_CASE(256) AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096)
#undef AE_CASE
    default: note_unsupported_launch("fused chain: FFT length must be a power of two in 256..4096");
  }
}

// -------------------------------------------------------------------------------------------------
// K12 OFDM-like chain (BASELINE config 5).  One frame per frame slot:
//   bits   : LTE x1 M-sequence x[n] = (x[n-28] + x[n-31]) % 2 (src/sequence.rs:42) seeded with
//            expand(frame_id + 1, 31) (src/sequence.rs:18-21), 2N bits per frame; the LFSR is linear, so
//            each thread XORs the host-tabulated words of the unit-seed sequences selected by the seed bits.
//   symbols: QPSK table (src/modulation.rs:87-92), idx = (b1<<1)+b0
//   tx     : Cfft::bwd with Scale::SN            (src/fft.rs:173-182)
//   channel: Awgn::apply semantics               (src/noise.rs:53-59), Philox stream = frame id,
//            block index = position mod N/2, words {0,1} for the lower and {2,3} for the upper half
//   rx     : Cfft::fwd with Scale::SN, demod_naive, bit errors and EVM partial sums.
// Compulsory HBM traffic: only the audit bits (4 B/symbol) when requested.
// -------------------------------------------------------------------------------------------------
constexpr uint64_t kLteX1PolyLow = (1ull << 0) | (1ull << 3);  // p(z) = z^31 + z^3 + 1  (back offsets 31, 28)

template <int N>
struct OfdmLaunch {
  static constexpr int T = FftCfg<N>::T;
  static constexpr int F = T >= 128 ? 1 : (128 / T);
  static constexpr int THREADS = F * T;
  static constexpr int WORDS = 2 * N / 32;  // packed bits per frame
  static constexpr size_t SMEM_PER_FRAME = (size_t)FftCfg<N>::SMEM_ELEMS * sizeof(float2) + WORDS * sizeof(uint32_t);
};

template <int N, bool FWD_INV>  // FWD_INV: exponent sign of Cfft::fwd is + (compat=reference)
__global__ void __launch_bounds__(OfdmLaunch<N>::THREADS, OfdmLaunch<N>::THREADS <= 128 ? 4 : 2)
ofdm_chain_kernel(size_t frames, uint64_t first_frame, float noise_scale, int twice, const __grid_constant__ PhiloxKeys keys, const float2* __restrict__ tw,
                  const uint32_t* __restrict__ zcol, int compat, uint8_t* __restrict__ tx_bits, uint8_t* __restrict__ rx_bits,
                  ae_stats* stats) {
  using C = FftCfg<N>;
  using LC = OfdmLaunch<N>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f = threadIdx.x / C::T;
  const int t = threadIdx.x % C::T;
  unsigned char* my = smem_raw + (size_t)f * LC::SMEM_PER_FRAME;
  float2* sm = reinterpret_cast<float2*>(my);
  uint32_t* words = reinterpret_cast<uint32_t*>(my + (size_t)C::SMEM_ELEMS * sizeof(float2));
  // QPSK table (src/modulation.rs:87-92) 0:(1,1) 1:(-1,1) 2:(1,-1) 3:(-1,-1): re = +-1 by bit 0, im = +-1 by
  // bit 1 of the index, so the symbol is 1.0f with the index bits moved into the sign bits (an indexed
  // local table compiled into four compares and eight predicated moves per symbol)
  auto qpsk_symbol = [](unsigned two) {
    return make_float2(__uint_as_float(0x3f800000u | ((two & 1u) << 31)), __uint_as_float(0x3f800000u | ((two >> 1) << 31)));
  };
  const float sn = 1.0f / sqrtf((float)N);  // Scale::SN (src/fft.rs:26)
  const unsigned hi_shift = compat == AE_COMPAT_REFERENCE ? 9u : 8u;
  unsigned long long errs = 0;
  double e_sum = 0.0, r_sum = 0.0;
  // persistent frame slots (see chain_fused_kernel)
  for (size_t fl = (size_t)blockIdx.x * LC::F + f; fl < frames; fl += (size_t)gridDim.x * LC::F) {
    const uint64_t frame_id = first_frame + fl;
    // ---- M-sequence: WORDS = N/16 = T, exactly one 32-bit word per thread.  The LFSR is linear over
    //      GF(2): word t of the sequence started from `state` is the XOR, over the set bits i of the
    //      state, of word t of the sequence started from the unit state e_i (host table zcol[i][t]) ----
    {
      const uint32_t state = (uint32_t)((frame_id + 1) & 0x7fffffffull);  // expand(frame_id+1, 31)
      uint32_t w = 0;
      // one predicated, independent load per seed bit (the bit test is uniform across the slot); a
      // find-first-set loop chained every table read behind the previous one (11 % of the stall samples)
#pragma unroll
      for (int i = 0; i < 31; ++i)
        if ((state >> i) & 1u) w ^= __ldg(zcol + i * C::T + t);
      words[t] = w;
    }
    frame_sync<C::T>(f);

    float2 v[16];
    uint32_t txb = 0;  // 2 bits per owned symbol
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      const int pos = t + m * C::T;
      const uint32_t w = words[pos >> 4];
      const unsigned two = (w >> ((2 * pos) & 31)) & 3u;  // bit0 = b0, bit1 = b1 -> idx = (b1<<1)+b0
      txb |= two << (2 * m);
      v[m] = qpsk_symbol(two);
    }
    if (tx_bits) {
      uint16_t* o = reinterpret_cast<uint16_t*>(tx_bits + 2 * fl * (size_t)N);
#pragma unroll
      for (int m = 0; m < 16; ++m) {
        const unsigned two = (txb >> (2 * m)) & 3u;
        o[t + m * C::T] = (uint16_t)((two & 1u) | ((two >> 1) << 8));
      }
    }
    // ---- tx: Cfft::bwd, Scale::SN ----
    fft_frame<N, !FWD_INV>(v, sm, tw, t, f);
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = cx_scale_exact(v[m], sn);
    // ---- channel ----
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      float2 z0, z1;
      awgn_unit_pair(keys, frame_id, (uint64_t)(t + m * C::T), z0, z1);
      z0 = cx_scale_exact(z0, noise_scale); z1 = cx_scale_exact(z1, noise_scale);
      if (twice) { z0 = cx_scale_exact(z0, noise_scale); z1 = cx_scale_exact(z1, noise_scale); }
      v[m] = cx_add_exact(v[m], z0);
      v[m + 8] = cx_add_exact(v[m + 8], z1);
    }
    // ---- rx: Cfft::fwd, Scale::SN, demod ----
    frame_sync<C::T>(f);
    fft_frame<N, FWD_INV>(v, sm, tw, t, f);
    float e_pow = 0.0f;
    unsigned rxb = 0;
    // the sign test is the reference's decision when min |component| > 2^-21 (1 + max |component|)^2 (common.cuh); the bound
    // grows with the maximum, so ONE test per thread on the minimum and the NaN-propagating maximum of its 32 components
    // replaces sixteen per-symbol tests (FMNMX3: two new values per instruction)
    float mx = 0.0f, mn = 3.0e38f;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      v[m] = cx_scale_exact(v[m], sn);
      const unsigned two = (txb >> (2 * m)) & 3u;
      asm("max.NaN.f32 %0, %0, %1, %2;" : "+f"(mx) : "f"(fabsf(v[m].x)), "f"(fabsf(v[m].y)));
      asm("min.f32 %0, %0, %1, %2;" : "+f"(mn) : "f"(fabsf(v[m].x)), "f"(fabsf(v[m].y)));
      rxb |= ((__float_as_uint(v[m].x) >> 31) | ((__float_as_uint(v[m].y) >> 31) << 1)) << (2 * m);
      const float2 ref = qpsk_symbol(two);
      const float dr = v[m].x - ref.x, di = v[m].y - ref.y;
      e_pow += dr * dr + di * di;
    }
    const float uu = fmaf(mx, 6.9053396600248786e-4f, 6.9053396600248786e-4f);   // 2^-10.5 (1 + max)
    if (!(mn > uu * uu)) {             // rare: a component near an axis, NaN or inf somewhere in this thread's symbols
#pragma unroll 1
      for (int m = 0; m < 16; ++m) {
        float2 sv = v[0];
#pragma unroll
        for (int k = 1; k < 16; ++k) sv = (m == k) ? v[k] : sv;       // select without dynamic register indexing
        if (!qpsk_fast_ok(sv)) rxb = (rxb & ~(3u << (2 * m))) | (demod_qpsk_exact_slow(sv) << (2 * m));
      }
    }
    errs += __popc(rxb ^ txb);
    e_sum += (double)e_pow;
    r_sum += 32.0;  // 16 symbols of power 2
    if (rx_bits) {
      uint16_t* ro = reinterpret_cast<uint16_t*>(rx_bits + 2 * fl * (size_t)N);
#pragma unroll
      for (int m = 0; m < 16; ++m) ro[t + m * C::T] = (uint16_t)qpsk_pair_from_index((rxb >> (2 * m)) & 3u, hi_shift);
    }
    frame_sync<C::T>(f);  // words[] and sm are rewritten by the next frame
  }
  if (stats) {
    unsigned long long nb = (errs || e_sum > 0.0 || r_sum > 0.0) ? (unsigned long long)(r_sum) : 0ull;  // 32 per frame = 2 bits x 16 symbols
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      errs += __shfl_xor_sync(0xffffffffu, errs, o);
      nb += __shfl_xor_sync(0xffffffffu, nb, o);
      e_sum += __shfl_xor_sync(0xffffffffu, e_sum, o);
      r_sum += __shfl_xor_sync(0xffffffffu, r_sum, o);
    }
    if ((threadIdx.x & 31) == 0 && nb) {
      atomicAdd(reinterpret_cast<unsigned long long*>(&stats->bit_errors), errs);
      atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_bits), nb);
      atomicAdd(&stats->err_pow, e_sum);
      atomicAdd(&stats->ref_pow, r_sum);
    }
  }
}

bool ofdm_supported(size_t nfft) { return nfft >= 512 && nfft <= 4096 && (nfft & (nfft - 1)) == 0; }

// host: zcol[i][t] = bits [32t, 32t+32) of the LTE x1 sequence x[n] = x[n-28] ^ x[n-31] started from the
// unit state e_i (x[i] = 1, the other 30 seed bits 0), for i < 31 and t < T (= N/16 words = 2N bits)
static const uint32_t* ofdm_column_table(int T, cudaStream_t st) {
  static std::map<std::pair<int, int>, uint32_t*> cache;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(dev, T);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  const size_t nbits = (size_t)T * 32;
  std::vector<uint32_t> h((size_t)31 * T, 0u);
  std::vector<uint8_t> x(nbits);
  for (int i = 0; i < 31; ++i) {
    for (size_t n = 0; n < 31 && n < nbits; ++n) x[n] = (n == (size_t)i);
    for (size_t n = 31; n < nbits; ++n) x[n] = x[n - 28] ^ x[n - 31];
    for (size_t n = 0; n < nbits; ++n)
      if (x[n]) h[(size_t)i * T + n / 32] |= 1u << (n % 32);
  }
  uint32_t* d = nullptr;
  // on failure the CUDA error stays pending and is reported by the caller's launch check
  if (cudaMalloc((void**)&d, h.size() * sizeof(uint32_t)) != cudaSuccess) return nullptr;
  if (cudaMemcpyAsync(d, h.data(), h.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  cache[key] = d;
  return d;
}

template <int N>
static void launch_ofdm_n(size_t frames, uint64_t first_frame, float noise_scale, int twice, uint64_t seed, const float2* tw,
                          int compat, uint8_t* tx_bits, uint8_t* rx_bits, ae_stats* stats, cudaStream_t st) {
  using LC = OfdmLaunch<N>;
  const size_t smem = LC::SMEM_PER_FRAME * LC::F;
  const size_t want = (frames + LC::F - 1) / LC::F;
  const uint32_t* zj = ofdm_column_table(FftCfg<N>::T, st);
  if (!zj) return;
  auto launch = [&](auto kern) {
    const size_t resident = resident_ctas((const void*)kern, LC::THREADS, smem);
    const unsigned grid = (unsigned)(want < resident ? want : resident);
    kern<<<grid, LC::THREADS, smem, st>>>(frames, first_frame, noise_scale, twice, make_philox_keys(seed), tw, zj, compat, tx_bits, rx_bits, stats);
  };
  if (compat == AE_COMPAT_REFERENCE) launch(ofdm_chain_kernel<N, true>);
  else launch(ofdm_chain_kernel<N, false>);
}

void launch_ofdm_chain(size_t nfft, size_t frames, uint64_t first_frame, float noise_scale, int twice, uint64_t seed,
                       const float2* tw, int compat, uint8_t* tx_bits, uint8_t* rx_bits, ae_stats* stats, cudaStream_t st) {
  if (frames == 0) return;
  switch (nfft) {
#define AE_CASE(NN) case NN: launch_ofdm_n<NN>(frames, first_frame, noise_scale, twice, seed, tw, compat, tx_bits, rx_bits, stats, st); break;
    AE_CASE(512) AE_CASE(1024) AE_CASE(2048) AE_CASE(4096)
#undef AE_CASE
    default: note_unsupported_launch("OFDM chain: FFT length must be a power of two in 512..4096");
  }
}

}  // namespace ae
