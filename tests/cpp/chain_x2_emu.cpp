// chain_x2_emu.cpp — runs the DEVICE code of K14b (aether_primitives_b200/csrc/chain_x2.cuh) on the host:
// one std::thread per CUDA thread, pthread barriers for __syncwarp / __syncthreads, a model of
// mma.sync.m16n8k8 (TF32 inputs, fragment layout of the PTX ISA) and of the TMA bulk copy + mbarrier.
// tests/test_chain_x2_emu.py feeds it frames and compares the decision bytes with the oracle, so
// the kernel's index maps, twiddle rows and fragment layouts are checked without a GPU.
//
// usage: chain_x2_emu <in.bin> <out.bin>
//   in.bin : int32 header {nfft, ntaps, frames, compat, inverse, staged (bit 0) | lean (bit 1), warps, blocks}, float32 scale,
//            frames*nfft cf32 samples, ntaps cf32 taps
//   out.bin: 2*frames*nfft decision bytes
#define AE_HOST_EMU 1
#define __device__
#define __global__
#define __host__
#define __shared__
#define __grid_constant__
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#include <pthread.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include <vector_functions.h>
#include <vector_types.h>

// ---- the intrinsics common.cuh / fft_device.cuh mention -------------------------------------------
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fma(a, b, c); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __uint2float_rn(uint32_t u) { return (float)u; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static const struct { unsigned x, y, z; } blockDim = {0, 0, 0};

#include "../../aether_primitives_b200/csrc/chain_x2_host.h"

// ---- execution model ----------------------------------------------------------------------------------
struct WarpCtx {
  pthread_barrier_t bar;
  uint32_t a[32][4], b[32][2];
};
struct CtaCtx {
  pthread_barrier_t bar;
  std::vector<WarpCtx> warps;
};
static thread_local WarpCtx* tl_warp = nullptr;
static thread_local CtaCtx* tl_cta = nullptr;
static thread_local int tl_lane = 0;

namespace ae {
void emu_syncwarp() { pthread_barrier_wait(&tl_warp->bar); }
void emu_syncthreads() { pthread_barrier_wait(&tl_cta->bar); }
static inline float tf32(uint32_t v) { return __uint_as_float(v & 0xffffe000u); }
// D(16x8) += A(16x8, row major) * B(8x8, column major); PTX ISA "Matrix Fragments for mma.m16n8k8":
//   a0:(g, tig) a1:(g+8, tig) a2:(g, tig+4) a3:(g+8, tig+4);  b0:(k=tig, n=g) b1:(k=tig+4, n=g);
//   d0:(g, 2tig) d1:(g, 2tig+1) d2:(g+8, 2tig) d3:(g+8, 2tig+1),  g = lane/4, tig = lane%4
void emu_mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  WarpCtx* w = tl_warp;
  for (int i = 0; i < 4; ++i) w->a[tl_lane][i] = a[i];
  for (int i = 0; i < 2; ++i) w->b[tl_lane][i] = b[i];
  pthread_barrier_wait(&w->bar);
  const int g = tl_lane >> 2, tig = tl_lane & 3;
  auto A = [&](int row, int col) { return tf32(w->a[(row & 7) * 4 + (col & 3)][(row >> 3) + 2 * (col >> 2)]); };
  auto B = [&](int k, int n) { return tf32(w->b[n * 4 + (k & 3)][k >> 2]); };
  for (int i = 0; i < 4; ++i) {
    const int row = g + 8 * (i >> 1), col = 2 * tig + (i & 1);
    float acc = d[i];
    for (int k = 0; k < 8; ++k) acc = std::fma(A(row, k), B(k, col), acc);
    d[i] = acc;
  }
  pthread_barrier_wait(&w->bar);
}
void emu_tma_issue(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  std::memcpy(dst_smem, src, bytes);
  reinterpret_cast<std::atomic<uint64_t>*>(bar)->fetch_add(1, std::memory_order_release);
}
void emu_tma_wait(uint64_t* bar, uint32_t phase) {
  while ((reinterpret_cast<std::atomic<uint64_t>*>(bar)->load(std::memory_order_acquire) & 1u) == phase) std::this_thread::yield();
}
}  // namespace ae

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror("in"); return 2; }
  int32_t hdr[8];
  float scale;
  if (std::fread(hdr, 4, 8, f) != 8 || std::fread(&scale, 4, 1, f) != 1) return 2;
  const int nfft = hdr[0], ntaps = hdr[1], frames = hdr[2], compat = hdr[3], inverse = hdr[4], staged = hdr[5], warps = hdr[6], blocks = hdr[7];
  if (nfft != 1024 || ntaps < 1 || ntaps > 64) { std::fprintf(stderr, "unsupported shape\n"); return 2; }
  std::vector<float2> x((size_t)frames * nfft), taps(ntaps);
  if (std::fread(x.data(), 8, x.size(), f) != x.size() || std::fread(taps.data(), 8, taps.size(), f) != taps.size()) return 2;
  std::fclose(f);

  // plan, as ae_chain_create builds it
  const double sgn = inverse ? +1.0 : -1.0;
  std::vector<float2> window(nfft), tw, hi, lo;
  for (int m = 0; m < nfft; ++m) {
    double ar = 0, ai = 0;
    for (int k = 0; k < ntaps; ++k) {
      const double a = -sgn * 2.0 * M_PI * (double)(((long long)m * k) % nfft) / (double)nfft;
      const double cr = std::cos(a), ci = std::sin(a);
      ar += taps[k].x * cr - taps[k].y * ci;
      ai += taps[k].x * ci + taps[k].y * cr;
    }
    window[m] = make_float2((float)(ar * (double)scale), (float)(ai * (double)scale));
  }
  ae::chain_x2_twiddles(nfft, tw);
  ae::chain_x2_split_taps(taps.data(), ntaps, hi, lo);
  std::vector<uint8_t> bits((size_t)2 * frames * nfft, 0xEE);

  ae::ChainX2Params p;
  p.x = x.data(); p.bits = bits.data(); p.frames = (size_t)frames; p.window = window.data(); p.tw = tw.data();
  p.taps_hi = hi.data(); p.taps_lo = lo.data(); p.ntaps = ntaps; p.scale = scale; p.compat = compat; p.debug = 0; p.stagger = 0;

  using XC = ae::X2Cfg<1024>;
  const size_t smem = XC::smem_bytes(warps, (staged & 1) != 0);
  for (int b = 0; b < blocks; ++b) {
    std::vector<unsigned char> sm(smem + 16);
    unsigned char* base = sm.data() + ((16 - ((uintptr_t)sm.data() & 15)) & 15);
    CtaCtx cta;
    cta.warps = std::vector<WarpCtx>(warps);
    pthread_barrier_init(&cta.bar, nullptr, 32 * warps);
    for (auto& w : cta.warps) pthread_barrier_init(&w.bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int tid = 0; tid < 32 * warps; ++tid) {
      th.emplace_back([&, tid]() {
        tl_cta = &cta; tl_warp = &cta.warps[tid >> 5]; tl_lane = tid & 31;
        const ae::X2Launch L{tid, b, blocks, 32 * warps};
        const bool stg = (staged & 1) != 0, lean = (staged & 2) != 0;   // bit 1: LEAN build (twiddles from shared memory)
        if (inverse) {
          if (stg) ae::chain_x2_body<1024, true, true, false>(p, L, base);
          else if (lean) ae::chain_x2_body<1024, true, false, true>(p, L, base);
          else ae::chain_x2_body<1024, true, false, false>(p, L, base);
        } else {
          if (stg) ae::chain_x2_body<1024, false, true, false>(p, L, base);
          else if (lean) ae::chain_x2_body<1024, false, false, true>(p, L, base);
          else ae::chain_x2_body<1024, false, false, false>(p, L, base);
        }
      });
    }
    for (auto& t : th) t.join();
  }
  f = std::fopen(argv[2], "wb");
  if (!f) { std::perror("out"); return 2; }
  std::fwrite(bits.data(), 1, bits.size(), f);
  std::fclose(f);
  return 0;
}
