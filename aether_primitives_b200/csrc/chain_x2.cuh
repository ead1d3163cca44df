// chain_x2.cuh — K14b: the headline chain (Cfft::fwd(scale) -> zero-state FIR -> QPSK demod_naive) for
// N = 1024 and up to 64 taps, one WARP per frame, built around the two things ncu showed to bind K14
// (chain.cu): FP32 issue slots and the FMA pipe.
//
// Same algorithm as K14 (see chain.cu): B = DFT(x .* w) is the circular convolution of X = DFT(x) with
// the taps, A = DFT(x) supplies the last T-1 bins for the wrap-around correction
//     y[n] = B[n] - fix[n],   fix[n] = sum_{i=0}^{T-2-n} h[n+1+i] rt[i],   rt[i] = scale * A[N-1-i].
// What is different:
//   * packed FP32 (SASS FFMA2/FADD2/FMUL2).  A first radix-2 DIF stage splits the frame into the
//     sub-transform of the EVEN bins, e[n] = x[n] + x[n+N/2], and of the ODD bins,
//     o[n] = (x[n] - x[n+N/2]) W_N^n; both are N/2-point transforms with identical twiddles, so they run
//     as the two halves of f32x2 registers: every butterfly instruction does two transforms' work, the
//     twiddle is a broadcast scalar operand.  The scalar first stage writes straight into the register
//     halves, so no instruction is spent on packing.
//   * 1024 = 32 * 32 with ONE exchange.  Thread t first transforms its 32 samples x[t + 32 j] over j (the
//     scalar radix-2 split above followed by a packed radix-16: register k0 then holds Y[2 k0] and
//     Y[2 k0 + 1] in its two halves), multiplies by W^(t kappa) and writes them to shared memory; thread
//     kappa reads the 32 values Z_t[kappa] back as 16 pairs (t even, t odd), runs a packed radix-16 over
//     the pairs and combines the two halves with a scalar radix-2 (compile-time twiddles W32^c):
//     X[kappa + 32 c].  The first build (512 = 16*16*2 per half, two exchanges) was bound by shared-memory
//     wavefronts (ncu: l1tex data pipe 77 %, 791 wavefronts per frame); this layout needs 144 fewer.
//     A frame belongs to one warp: the exchange is __syncwarp()-only, warps drift out of phase.
//   * every twiddle a thread needs is fixed (it depends on the lane only): loaded once into registers.
//   * the wrap-around correction (T(T-1)/2 complex MACs per frame, 20 % of K14's FMA-pipe time) runs on
//     the tensor cores: with n = 8a + b it is the GEMM
//         fix[b][a] = sum_{i'} H[b][i'] R[i'][a],   H[b][i'] = h[b+1+i'],   R[i'][a] = rt[i' - 8a],
//     M = 16 (b, real and imaginary tap parts), N = 8 (a), K = 64, as mma.sync m16n8k8 TF32 with the
//     3xTF32 split (hi*hi + hi*lo + lo*hi: relative error ~2^-21, FP32 class).  48 MMAs per frame.
//   * thread kappa ends with bins kappa + 32 c: the two decision bytes of a bin are one 16-bit store,
//     a warp writes 64 contiguous bytes per instruction.
//
// The file is also compiled for the HOST by tests/cpp/chain_x2_emu.cpp (one std::thread per CUDA
// thread, barriers for __syncwarp/__syncthreads, an m16n8k8 emulation) so that index maps, twiddle
// rows and fragment layouts are checked against the oracle without a GPU.
#pragma once
#include "../../include/aether_b200.h"
#ifndef AE_HOST_EMU
#include "async_copy.cuh"
#endif
#include "fft_device.cuh"

namespace ae {

// ---- environment: device intrinsics or their host emulation ---------------------------------------
#ifdef AE_HOST_EMU
// provided by the emulation harness
void emu_syncwarp();
void emu_syncthreads();
void emu_mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]);
void emu_tma_issue(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar);
void emu_tma_wait(uint64_t* bar, uint32_t phase);
#define AE_X2_DEV
static inline float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
static inline float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
static inline float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
static inline float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)); }
static inline void x2_syncwarp() { emu_syncwarp(); }
static inline void x2_syncthreads() { emu_syncthreads(); }
static inline void x2_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) { emu_mma_tf32(d, a, b); }
static inline float x2_fmax3_nan(float a, float b, float c) {
  if (a != a || b != b || c != c) return NAN;
  return std::fmax(a, std::fmax(b, c));
}
static inline float x2_fmin3(float a, float b, float c) { return std::fmin(a, std::fmin(b, c)); }
static inline uint32_t x2_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t d = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t nib = (sel >> (4 * i)) & 0xfu;
    uint32_t byte = (uint32_t)(src >> (8 * (nib & 7u))) & 0xffu;
    if (nib & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
    d |= byte << (8 * i);
  }
  return d;
}
#else
#define AE_X2_DEV __device__ __forceinline__
AE_X2_DEV float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
AE_X2_DEV float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
AE_X2_DEV float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
AE_X2_DEV float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
AE_X2_DEV void x2_syncwarp() { __syncwarp(); }
AE_X2_DEV void x2_syncthreads() { __syncthreads(); }
// D += A(16x8, row) * B(8x8, col), TF32 inputs, FP32 accumulate (SASS HMMA.1688.F32.TF32)
AE_X2_DEV void x2_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// NaN-propagating three-input maximum / plain three-input minimum (sm_100: FMNMX3)
AE_X2_DEV float x2_fmax3_nan(float a, float b, float c) {
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
AE_X2_DEV float x2_fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// byte permute; a selector nibble with bit 3 set replicates the SIGN of the selected byte
AE_X2_DEV uint32_t x2_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
#endif

AE_X2_DEV float2 f2bc(float s) { return make_float2(s, s); }

// two complex values that share every operation: (re.x, im.x) and (re.y, im.y)
struct cx2 { float2 re, im; };
AE_X2_DEV cx2 c2add(cx2 a, cx2 b) { return cx2{f2add(a.re, b.re), f2add(a.im, b.im)}; }
AE_X2_DEV cx2 c2sub(cx2 a, cx2 b) { return cx2{f2sub(a.re, b.re), f2sub(a.im, b.im)}; }
// a * (c - i s) forward, a * (c + i s) inverse; c, s scalars (immediates when known at compile time)
template <bool INV>
AE_X2_DEV cx2 c2mul_w(cx2 a, float c, float s) {
  if (INV) return cx2{f2fma(a.im, f2bc(-s), f2mul(a.re, f2bc(c))), f2fma(a.re, f2bc(s), f2mul(a.im, f2bc(c)))};
  return cx2{f2fma(a.im, f2bc(s), f2mul(a.re, f2bc(c))), f2fma(a.re, f2bc(-s), f2mul(a.im, f2bc(c)))};
}
// table twiddle w = exp(-i theta): a * w forward, a * conj(w) inverse
template <bool INV>
AE_X2_DEV cx2 c2mul_tw(cx2 a, float2 w) { return c2mul_w<INV>(a, w.x, -w.y); }

// 4-point DFT; the +-i rotation is folded into which halves are added (no negation is ever issued)
template <bool INV>
AE_X2_DEV void c2dft4(cx2& a0, cx2& a1, cx2& a2, cx2& a3) {
  const cx2 s0 = c2add(a0, a2), s1 = c2sub(a0, a2), s2 = c2add(a1, a3), d = c2sub(a1, a3);
  a0 = c2add(s0, s2);
  a2 = c2sub(s0, s2);
  if (INV) {  // s3 = +i d = (-d.im, d.re)
    a1 = cx2{f2sub(s1.re, d.im), f2add(s1.im, d.re)};
    a3 = cx2{f2add(s1.re, d.im), f2sub(s1.im, d.re)};
  } else {    // s3 = -i d = (d.im, -d.re)
    a1 = cx2{f2add(s1.re, d.im), f2sub(s1.im, d.re)};
    a3 = cx2{f2sub(s1.re, d.im), f2add(s1.im, d.re)};
  }
}
// same with a2 still to be multiplied by -i (forward) / +i (inverse): W16^4 of the radix-16 butterfly
template <bool INV>
AE_X2_DEV void c2dft4_rot2(cx2& a0, cx2& a1, cx2& a2, cx2& a3) {
  cx2 s0, s1;
  if (INV) {  // a2' = (-a2.im, a2.re)
    s0 = cx2{f2sub(a0.re, a2.im), f2add(a0.im, a2.re)};
    s1 = cx2{f2add(a0.re, a2.im), f2sub(a0.im, a2.re)};
  } else {    // a2' = (a2.im, -a2.re)
    s0 = cx2{f2add(a0.re, a2.im), f2sub(a0.im, a2.re)};
    s1 = cx2{f2sub(a0.re, a2.im), f2add(a0.im, a2.re)};
  }
  const cx2 s2 = c2add(a1, a3), d = c2sub(a1, a3);
  a0 = c2add(s0, s2);
  a2 = c2sub(s0, s2);
  if (INV) {
    a1 = cx2{f2sub(s1.re, d.im), f2add(s1.im, d.re)};
    a3 = cx2{f2add(s1.re, d.im), f2sub(s1.im, d.re)};
  } else {
    a1 = cx2{f2add(s1.re, d.im), f2sub(s1.im, d.re)};
    a3 = cx2{f2sub(s1.re, d.im), f2add(s1.im, d.re)};
  }
}
template <int E, bool INV>
AE_X2_DEV cx2 c2mul_w16(cx2 a) {  // a * W16^E (E not a multiple of 4), see mul_w16 in fft_device.cuh
  constexpr int e = ((E % 16) + 16) % 16;
  static_assert(e % 4 != 0, "rotations by multiples of -i are folded into the butterflies");
  constexpr float c = (e == 1 || e == 15) ? kC16 : (e == 2 || e == 14) ? kC8 : (e == 3 || e == 13) ? kS16
                      : (e == 5 || e == 11) ? -kS16 : (e == 6 || e == 10) ? -kC8 : -kC16;
  constexpr float s = (e == 1 || e == 7) ? kS16 : (e == 2 || e == 6) ? kC8 : (e == 3 || e == 5) ? kC16
                      : (e == 9 || e == 15) ? -kS16 : (e == 10 || e == 14) ? -kC8 : -kC16;
  return c2mul_w<INV>(a, c, s);
}
// natural-order 16-point DFT (n = 4a + b, k = c + 4d), as Dft<16> in fft_device.cuh
template <bool INV>
AE_X2_DEV void c2dft16(cx2 (&v)[16]) {
#pragma unroll
  for (int b = 0; b < 4; ++b) c2dft4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]);
  v[5] = c2mul_w16<1, INV>(v[5]);
  v[6] = c2mul_w16<2, INV>(v[6]);
  v[7] = c2mul_w16<3, INV>(v[7]);
  v[9] = c2mul_w16<2, INV>(v[9]);
  v[11] = c2mul_w16<6, INV>(v[11]);
  v[13] = c2mul_w16<3, INV>(v[13]);
  v[14] = c2mul_w16<6, INV>(v[14]);
  v[15] = c2mul_w16<9, INV>(v[15]);
  c2dft4<INV>(v[0], v[1], v[2], v[3]);
  c2dft4<INV>(v[4], v[5], v[6], v[7]);
  c2dft4_rot2<INV>(v[8], v[9], v[10], v[11]);   // v[10] * W16^4
  c2dft4<INV>(v[12], v[13], v[14], v[15]);
  cx2 o[16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = 0; d < 4; ++d) o[c + 4 * d] = v[4 * c + d];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = o[i];
}

// ---- geometry ---------------------------------------------------------------------------------------
template <int N>
struct X2Cfg {
  static constexpr int T = 32;                   // threads per frame: one warp
  static_assert(N == 1024, "K14b is laid out for N = 1024 = 32 * 32: one warp per frame");
  static constexpr int MAX_TAPS = 64;            // wrap-around correction needs the last 64 bins only (c = 30, 31)
  // per-thread twiddle rows (row-major [row][t]): 16 first-stage rows W_N^(t + 32 m), 6 rows W_512^(e t), e = 1,2,3,4,8,12
  static constexpr int ROW_S1 = 0, ROW_P1 = 16, ROWS = 22;
  static constexpr int HPAD = MAX_TAPS + 8;      // zero padded tap arrays (hi / lo parts)
  // shared memory (bytes): [taps hi: HPAD cf32][taps lo: HPAD cf32][window: N cf32][first-stage twiddles: 16 T cf32]
  //                        [per warp: ex | rtz | fix | xin | mbar]
  // exchange buffer: two float planes (re, im); Z_t[kappa] at float kappa*36 + t.  Writers (lanes = t) store 4-byte
  // words to consecutive addresses; reader kappa fetches its row with eight 16-byte loads per plane (row stride 36
  // floats = 9 x 16 B: the 8 lanes of a quarter-warp hit 8 different bank groups), and every 16-byte load is two
  // ready-made (t even, t odd) register pairs.  (Dispatch cost per frame, from the SASS stall fields: 64 LDS.32 + 29
  // MOVs = 450 cycles in the first one-exchange build, 32 LDS.128 now.)
  static constexpr int EX_ROW = 36;                          // floats per kappa row
  static constexpr int EX_ELEMS = 16 * EX_ROW;               // float2 per plane (= 32 rows of EX_ROW floats)
  // rt, pre-split for 3xTF32 and laid out as mma B fragments: block j = u >> 3 (u = 64 + i, i the rt index; blocks 0..7
  // stay zero) holds, for tig = (u & 3), the 8 floats {hi.re, lo.re, hi.im, lo.im} x {u & 4 ? 1 : 0} interleaved as
  // (hi.re[u], hi.re[u+4], lo.re[u], lo.re[u+4], hi.im[u], hi.im[u+4], lo.im[u], lo.im[u+4]): two 16-byte loads are the
  // four fragment register pairs of one k-step.  Block stride 36 floats: lanes g, g+1 fall into different bank groups.
  static constexpr int RTZ_BLOCK = 36;                       // floats per block of 8 rt entries (32 + 4 of skew)
  static constexpr int RTZ_ELEMS = 16 * RTZ_BLOCK / 2;       // float2
  static constexpr int FIX_ELEMS = 80;                       // n in [0,64): n + 4 (n >> 4)
  static constexpr size_t HEAD_BYTES = (size_t)(2 * HPAD + N + 16 * T) * sizeof(float2);   // + first-stage twiddle rows (LEAN)
  static constexpr size_t warp_bytes(bool staged) {
    return (size_t)2 * EX_ELEMS * sizeof(float2) + (size_t)(RTZ_ELEMS + FIX_ELEMS) * sizeof(float2) + (staged ? (size_t)N * sizeof(float2) : 0) + 16;
  }
  static constexpr size_t smem_bytes(int warps, bool staged) { return HEAD_BYTES + (size_t)warps * warp_bytes(staged); }
};
AE_X2_DEV int x2_rtz_phys(int u) { return u + 4 * (u >> 3); }
AE_X2_DEV int x2_fix_phys(int n) { return n + 4 * (n >> 4); }

struct ChainX2Params {
  const float2* x;        // frames * N input samples
  uint8_t* bits;          // 2 bytes per sample, 4-byte aligned
  size_t frames;
  const float2* window;   // N cf32: scale * sum_k h[k] exp(-sgn 2 pi i m k/N)
  const float2* tw;       // X2Cfg::ROWS rows of T per-thread twiddles
  const float2* taps_hi;  // HPAD cf32: taps truncated to TF32, zero padded
  const float2* taps_lo;  // HPAD cf32: taps - taps_hi
  int ntaps;
  float scale;
  int compat;
  int debug;              // developer switch for timing experiments (AE_CHAIN_DEBUG): 1 = skip transform A and the fix-up, 2 = skip B
  int stagger;            // cycles between the starts of the warps that share an SM sub-partition (see chain_x2_body)
};

// the lane's twiddles, loaded once
struct X2Tw {
  float2 s1[16];   // W_N^(t + 32 m)             first (radix-2, scalar) stage, odd half
  float2 wl[4];    // W_512^(b t), b = 1..3      between the stages: Z[2 k0 + b'] = W_512^(k0 t) Y[2 k0 + b'],
  float2 wh[4];    // W_512^(4 a t), a = 1..3    W^(4a+b) = W^(4a) W^b
};

// cos / sin of 2 pi c / 32, c = 0..15 (twiddles of the final radix-2 combine)
__host__ __device__ constexpr float x2_c32(int c) {
  constexpr float v[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f,
                           0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f, 0.0f, -0.19509032201612826785f,
                           -0.38268343236508977173f, -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                           -0.92387953251128675613f, -0.98078528040323044913f};
  return v[c];
}
__host__ __device__ constexpr float x2_s32(int c) { return c <= 8 ? x2_c32(8 - c) : x2_c32(c - 8); }   // sin t = cos(t - pi/2)

// Second half of the frame transform.  In: v[k0] = (Y[2 k0], Y[2 k0 + 1]) of thread t, Y = DFT_32 over j of x[t + 32 j]
// (the odd half already carries W_N^t).  Out: y[c] = X[kappa + 32 c] for thread kappa = t.
// TAIL: only y[30], y[31] are produced (the last 64 bins of the frame, all the wrap-around correction needs).
// LEAN: the nine twiddle products are recomputed per transform (an opaque move keeps the compiler from hoisting them
// out of the frame loop into 18 more registers).
AE_X2_DEV float2 x2_opaque(float2 w) {
#ifndef AE_HOST_EMU
  asm volatile("" : "+f"(w.x), "+f"(w.y));
#endif
  return w;
}
template <int N, bool INV, bool TAIL, bool LEAN>
AE_X2_DEV void x2_second_stage(cx2 (&v)[16], float2 (&y)[32], float2* ex, const X2Tw& tw, int t) {
  using XC = X2Cfg<N>;
  constexpr int EX = XC::EX_ELEMS;
  // Z_t[2 k0 + b] = W_512^(k0 t) v[k0].b, stored at float (k0*33 + t)*2 + b of each plane
  {
    // twiddle and store element by element: the 4-byte stores are spaced by the multiplies instead of queueing up
    float* wr = reinterpret_cast<float*>(ex) + t;
    float* wi = wr + 2 * EX;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      if (r > 0) {
        const int a = r >> 2, bb = r & 3;
        float2 w;
        if (a == 0) w = tw.wl[bb];
        else if (bb == 0) w = tw.wh[a];
        else w = cx_mul(LEAN ? x2_opaque(tw.wh[a]) : tw.wh[a], tw.wl[bb]);
        v[r] = c2mul_tw<INV>(v[r], w);
      }
      wr[(2 * r) * XC::EX_ROW] = v[r].re.x; wr[(2 * r + 1) * XC::EX_ROW] = v[r].re.y;
      wi[(2 * r) * XC::EX_ROW] = v[r].im.x; wi[(2 * r + 1) * XC::EX_ROW] = v[r].im.y;
    }
  }
  x2_syncwarp();
  // thread kappa gathers its row Z_t[kappa], t = 0..31: pairs (t = 2u, 2u + 1) = halves of the packed registers
  {
    const float4* pr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ex) + t * XC::EX_ROW);
    const float4* pi = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ex) + 2 * EX + t * XC::EX_ROW);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 a = pr[q], b = pi[q];
      v[2 * q] = cx2{make_float2(a.x, a.y), make_float2(b.x, b.y)};
      v[2 * q + 1] = cx2{make_float2(a.z, a.w), make_float2(b.z, b.w)};
    }
  }
  // X[kappa + 32 c] = sum_t Z_t W_32^(t c) = E[c mod 16] + W_32^c O[c mod 16], E / O = DFT_16 over the even / odd t
  c2dft16<INV>(v);
  if (TAIL) {
    // c = 30, 31: E[c0] - W_32^c0 O[c0], c0 = 14, 15
    static_for<14, 16>([&](auto cc) {
      constexpr int c0 = decltype(cc)::value;
      const float2 o = mul_w<INV>(make_float2(v[c0].re.y, v[c0].im.y), x2_c32(c0), x2_s32(c0));
      y[c0 + 16] = make_float2(v[c0].re.x - o.x, v[c0].im.x - o.y);
    });
  } else {
    static_for<0, 16>([&](auto cc) {
      constexpr int c0 = decltype(cc)::value;
      const float2 ev = make_float2(v[c0].re.x, v[c0].im.x), od = make_float2(v[c0].re.y, v[c0].im.y);
      if constexpr (c0 == 0 || c0 == 8) {
        const float2 o = c0 == 0 ? od : mul_mi<INV>(od);
        y[c0] = make_float2(ev.x + o.x, ev.y + o.y);
        y[c0 + 16] = make_float2(ev.x - o.x, ev.y - o.y);
      } else {
        // y+ = E + W O as two FMA chains, y- = 2 E - y+: 6 operations instead of 8
        constexpr float c = x2_c32(c0), sn = INV ? -x2_s32(c0) : x2_s32(c0);   // W = c - i sn
        y[c0] = make_float2(fmaf(od.x, c, fmaf(od.y, sn, ev.x)), fmaf(od.y, c, fmaf(-od.x, sn, ev.y)));
        y[c0 + 16] = make_float2(fmaf(2.0f, ev.x, -y[c0].x), fmaf(2.0f, ev.y, -y[c0].y));
      }
    });
  }
}

// first stage: radix-2 DIF split of the frame into the even-bin (x half) and odd-bin (y half) inputs;
// WIN multiplies by the FIR window first
// LEAN reads the odd half's twiddles from the shared table s1tab[m][t] instead of 32 registers
template <int N, bool INV, bool WIN, bool LEAN>
AE_X2_DEV void x2_stage1(cx2 (&v)[16], const float2* __restrict__ src, const float2* __restrict__ win, const X2Tw& tw,
                         const float2* __restrict__ s1tab, int t) {
  constexpr int T = X2Cfg<N>::T;
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const int s = t + m * T;
    const float2 a = src[s], b = src[s + N / 2];
    float2 e, d;
    if (WIN) {
      // e = a wa + b wb, d = a wa - b wb = 2 (a wa) - e: 10 FMA-pipe operations instead of 12
      const float2 wa = win[s], wb = win[s + N / 2];
      const float2 pa = cx_mul(a, wa);
      e = make_float2(fmaf(b.x, wb.x, fmaf(-b.y, wb.y, pa.x)), fmaf(b.x, wb.y, fmaf(b.y, wb.x, pa.y)));
      d = make_float2(fmaf(2.0f, pa.x, -e.x), fmaf(2.0f, pa.y, -e.y));
    } else {
      e = make_float2(a.x + b.x, a.y + b.y);
      d = make_float2(a.x - b.x, a.y - b.y);
    }
    const float2 o = mul_tw<INV>(d, LEAN ? s1tab[s] : tw.s1[m]);
    v[m] = cx2{make_float2(e.x, o.x), make_float2(e.y, o.y)};
  }
}

// first stage from registers: y[j] = sample at position t + 32 j (the layout x2_second_stage leaves its results in)
template <bool INV>
AE_X2_DEV void x2_stage1_regs(cx2 (&v)[16], const float2 (&y)[32], const X2Tw& tw) {
#pragma unroll
  for (int m = 0; m < 16; ++m) {
    const float2 a = y[m], b = y[m + 16];
    const float2 e = make_float2(a.x + b.x, a.y + b.y), d = make_float2(a.x - b.x, a.y - b.y);
    const float2 o = mul_tw<INV>(d, tw.s1[m]);
    v[m] = cx2{make_float2(e.x, o.x), make_float2(e.y, o.y)};
  }
}
// the lane's twiddle registers from the per-thread table (chain_x2_twiddles)
template <int N>
AE_X2_DEV void x2_load_twiddles(X2Tw& tw, const float2* __restrict__ table, int t) {
  using XC = X2Cfg<N>;
#pragma unroll
  for (int m = 0; m < 16; ++m) tw.s1[m] = table[(XC::ROW_S1 + m) * XC::T + t];
#pragma unroll
  for (int i = 1; i < 4; ++i) { tw.wl[i] = table[(XC::ROW_P1 + i - 1) * XC::T + t]; tw.wh[i] = table[(XC::ROW_P1 + i + 2) * XC::T + t]; }
  tw.wl[0] = tw.wh[0] = make_float2(1.0f, 0.0f);
}

// The two decision bytes of one bin as a little-endian u16: byte = sign of re / im, the im byte is 2
// (compat=reference, idx & 2) or 1 (corrected); `mask` = 0x0201 / 0x0101
AE_X2_DEV uint32_t x2_sign_pair(float2 v, uint32_t mask) {
  return x2_prmt(__float_as_uint(v.x), __float_as_uint(v.y), 0x00FBu) & mask;   // bytes 0,1 <- sign(re), sign(im) replicated
}

// rt entry u, split for the 3xTF32 products (hi = the value truncated to TF32's 10 mantissa bits, lo = value - hi) and
// written into the fragment-ready layout described at X2Cfg::RTZ_BLOCK
template <int BLOCK>
AE_X2_DEV void x2_store_rt(float* rtz, int u, float2 a) {
  const float hx = __uint_as_float(__float_as_uint(a.x) & 0xffffe000u), hy = __uint_as_float(__float_as_uint(a.y) & 0xffffe000u);
  float* q = rtz + (u >> 3) * BLOCK + 8 * (u & 3) + ((u >> 2) & 1);
  q[0] = hx; q[2] = a.x - hx; q[4] = hy; q[6] = a.y - hy;
}

struct X2Launch { int tid, bid, nblocks, nthreads; };

template <int N, bool INV, bool STAGED, bool LEAN = false>
AE_X2_DEV void chain_x2_body(const ChainX2Params& p, const X2Launch& L, unsigned char* smem_raw) {
  using XC = X2Cfg<N>;
  constexpr int T = XC::T;
  const int warp = L.tid >> 5, t = L.tid & 31, nwarps = L.nthreads >> 5;
  float2* hhi = reinterpret_cast<float2*>(smem_raw);
  float2* hlo = hhi + XC::HPAD;
  float2* win = hlo + XC::HPAD;
  float2* s1tab = win + N;                                                        // [m][t] = W_N^(t + 32 m) = W_N^s
  unsigned char* mine = smem_raw + XC::HEAD_BYTES + (size_t)warp * XC::warp_bytes(STAGED);
  float2* ex = reinterpret_cast<float2*>(mine);                                  // two planes of EX_ELEMS float2
  float2* rtz = ex + 2 * XC::EX_ELEMS;
  float2* fixb = rtz + XC::RTZ_ELEMS;
  float2* xin = fixb + XC::FIX_ELEMS;                                           // STAGED only
  uint64_t* bar = reinterpret_cast<uint64_t*>(xin + (STAGED ? N : 0));

  for (int i = L.tid; i < XC::HPAD; i += L.nthreads) { hhi[i] = p.taps_hi[i]; hlo[i] = p.taps_lo[i]; }
  for (int i = L.tid; i < N; i += L.nthreads) win[i] = p.window[i];
  for (int i = L.tid; i < 16 * T; i += L.nthreads) s1tab[i] = p.tw[XC::ROW_S1 * T + i];
  for (int i = t; i < XC::RTZ_ELEMS; i += 32) rtz[i] = make_float2(0.0f, 0.0f);   // logical [0,64) stays zero for ever
  X2Tw tw;
#pragma unroll
  for (int m = 0; m < 16; ++m) tw.s1[m] = LEAN ? make_float2(0.0f, 0.0f) : p.tw[(XC::ROW_S1 + m) * T + t];
#pragma unroll
  for (int i = 1; i < 4; ++i) { tw.wl[i] = p.tw[(XC::ROW_P1 + i - 1) * T + t]; tw.wh[i] = p.tw[(XC::ROW_P1 + i + 2) * T + t]; }
  tw.wl[0] = tw.wh[0] = make_float2(1.0f, 0.0f);

  const size_t stride = (size_t)L.nblocks * nwarps;
  size_t frame = (size_t)L.bid * nwarps + warp;
  if (STAGED) {
#ifdef AE_HOST_EMU
    if (t == 0) { *bar = 0; if (frame < p.frames) emu_tma_issue(xin, p.x + frame * N, N * (uint32_t)sizeof(float2), bar); }
#else
    if (t == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
      if (frame < p.frames) {
        mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
        bulk_g2s(xin, p.x + frame * N, N * (uint32_t)sizeof(float2), bar);
      }
    }
#endif
  }
  x2_syncthreads();
#ifndef AE_HOST_EMU
  // Stagger.  Every warp runs the same periodic mix of an FMA-pipe phase (butterflies) and a phase that
  // does not need that pipe (shared-memory exchange, decisions, MMA).  Warps that start together stay
  // together: the pipe is shared fairly while both want it, so their distance never changes, and the pipe
  // idles whenever they are all in the other phase (measured: throughput independent of the warp count,
  // FMA pipe 47 % busy).  Starting the warps of one sub-partition (warp id mod 4) a fraction of a frame
  // apart makes the phases of one warp fall into the gaps of the others.
  if (p.stagger > 0) {
    const long long wait = (long long)(warp >> 2) * p.stagger;
    const long long t0 = clock64();
    while (clock64() - t0 < wait) {}
  }
#endif

  const int g = t >> 2, tig = t & 3;                       // mma.sync fragment coordinates
  const int ksteps = (p.ntaps - 1 + 7) >> 3;               // k-steps of 8 that hold non-zero taps (<= 8)
  const uint32_t mask = p.compat == AE_COMPAT_REFERENCE ? 0x0201u : 0x0101u;
  const unsigned hi_shift = p.compat == AE_COMPAT_REFERENCE ? 9u : 8u;
  uint32_t phase = 0;
  for (; frame < p.frames; frame += stride) {
    const float2* src;
    if (STAGED) {
#ifdef AE_HOST_EMU
      emu_tma_wait(bar, phase);
#else
      mbar_wait(bar, phase);
#endif
      phase ^= 1u;
      src = xin;
    } else {
      src = p.x + frame * N;
#ifndef AE_HOST_EMU
      // plain loads have no look-ahead: pull this warp's next frame towards L2 (8 KB = 64 lines, two per lane)
      if (frame + stride < p.frames) {
        const char* pf = reinterpret_cast<const char*>(p.x + (frame + stride) * N) + 128 * t;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 4096));
      }
#endif
    }
    cx2 v[16];
    float2 y[32];
    float2 fx0 = make_float2(0.0f, 0.0f), fx1 = fx0;
#ifdef AE_CHAIN_DEBUG_BUILD
    if (!(p.debug & 1))
#endif
    {
    // ---- A = DFT(x), bins kappa + 32 c for c = 30, 31: the last 64 bins ----
    x2_stage1<N, INV, false, LEAN>(v, src, nullptr, tw, s1tab, t);
    c2dft16<INV>(v);
    x2_second_stage<N, INV, true, LEAN>(v, y, ex, tw, t);
    {
      // rt[i] = scale * A[N-1-i] at logical index u = 64 + i: y[31] = A[992 + t] -> i = 31 - t, y[30] = A[960 + t] -> i = 63 - t
      float2 a30 = cx_scale_exact(y[30], p.scale);
      const float2 a31 = cx_scale_exact(y[31], p.scale);
      if (t == 0) a30 = make_float2(0.0f, 0.0f);           // i = 63 is never used (taps beyond it are zero): keep 0 * inf out
      x2_store_rt<XC::RTZ_BLOCK>(reinterpret_cast<float*>(rtz), 64 + 31 - t, a31);
      x2_store_rt<XC::RTZ_BLOCK>(reinterpret_cast<float*>(rtz), 64 + 63 - t, a30);
    }
    x2_syncwarp();
    // ---- wrap-around correction on the tensor cores (3xTF32); six independent accumulator chains ----
    float d1[3][4], d2[3][4];   // H * Re(R), H * Im(R); rows 0-7: Re(H), rows 8-15: Im(H); [0] lo*hi, [1] hi*lo, [2] hi*hi
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) d1[i][j] = d2[i][j] = 0.0f;
    auto mma_step = [&](int ks) {
      const int hi0 = g + tig + 8 * ks + 1;                // A fragment: rows (g, g+8) = (Re h, Im h)[g + 1 + col], cols tig, tig + 4
      const float2 ah0 = hhi[hi0], ah1 = hhi[hi0 + 4], al0 = hlo[hi0], al1 = hlo[hi0 + 4];
      const uint32_t a_hi[4] = {__float_as_uint(ah0.x), __float_as_uint(ah0.y), __float_as_uint(ah1.x), __float_as_uint(ah1.y)};
      const uint32_t a_lo[4] = {__float_as_uint(al0.x), __float_as_uint(al0.y), __float_as_uint(al1.x), __float_as_uint(al1.y)};
      // B fragment: R[k][n] = rt[8 ks + k - 8 n], k = tig (+4), n = g  ->  block 8 + ks - g, slot tig
      const float4* rb = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rtz) + (8 + ks - g) * XC::RTZ_BLOCK + 8 * tig);
      const float4 r0 = rb[0], r1 = rb[1];
      const uint32_t bh[2] = {__float_as_uint(r0.x), __float_as_uint(r0.y)}, bl[2] = {__float_as_uint(r0.z), __float_as_uint(r0.w)};
      const uint32_t ch[2] = {__float_as_uint(r1.x), __float_as_uint(r1.y)}, cl[2] = {__float_as_uint(r1.z), __float_as_uint(r1.w)};
      x2_mma(d1[0], a_lo, bh); x2_mma(d1[1], a_hi, bl); x2_mma(d1[2], a_hi, bh);
      x2_mma(d2[0], a_lo, ch); x2_mma(d2[1], a_hi, cl); x2_mma(d2[2], a_hi, ch);
    };
    if (ksteps == 8) {                                     // 57..64 taps: straight-line code, no per-step branch
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) mma_step(ks);
    } else {
#pragma unroll 1
      for (int ks = 0; ks < ksteps; ++ks) mma_step(ks);
    }
    {
      // accumulator (row g, col 2 tig + j): Re(H) part for n = 8 (2 tig + j) + g; row g + 8: Im(H) part
      float e1[4], e2[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { e1[j] = (d1[0][j] + d1[1][j]) + d1[2][j]; e2[j] = (d2[0][j] + d2[1][j]) + d2[2][j]; }
      const int n1 = 16 * tig + g;
      fixb[x2_fix_phys(n1)] = make_float2(e1[0] - e2[2], e1[2] + e2[0]);
      fixb[x2_fix_phys(n1 + 8)] = make_float2(e1[1] - e2[3], e1[3] + e2[1]);
    }
    x2_syncwarp();
    fx0 = fixb[x2_fix_phys(t)]; fx1 = fixb[x2_fix_phys(32 + t)];   // bins t (c = 0) and 32 + t (c = 1)
    }
    // ---- B = DFT(x .* w): circular convolution of scale * X with the taps ----
    x2_stage1<N, INV, true, LEAN>(v, src, win, tw, s1tab, t);
    if (STAGED) {
      x2_syncwarp();                                       // every lane has consumed the staged frame
      if (t == 0 && frame + stride < p.frames) {
#ifdef AE_HOST_EMU
        emu_tma_issue(xin, p.x + (frame + stride) * N, N * (uint32_t)sizeof(float2), bar);
#else
        mbar_expect_tx(bar, N * (uint32_t)sizeof(float2));
        bulk_g2s(xin, p.x + (frame + stride) * N, N * (uint32_t)sizeof(float2), bar);
#endif
      }
    }
    c2dft16<INV>(v);
    x2_second_stage<N, INV, false, LEAN>(v, y, ex, tw, t);
    y[0].x -= fx0.x; y[0].y -= fx0.y;                      // bins < 64 are the only ones with wrap-around terms
    y[1].x -= fx1.x; y[1].y -= fx1.y;
    // ---- hard decisions (src/modulation.rs:33-56) ----
    // The sign test is the reference's answer whenever min(|re|,|im|) > 2^-21 (1 + max(|re|,|im|))^2
    // (common.cuh).  The bound is monotone in max, so one test per thread with the maximum and minimum
    // over all 64 of its components is sufficient; NaN propagates through the maximum and fails it.
    float mx = 0.0f, mn = 3.0e38f;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      mx = x2_fmax3_nan(mx, fabsf(y[c].x), fabsf(y[c].y));
      mn = x2_fmin3(mn, fabsf(y[c].x), fabsf(y[c].y));
    }
    const float uu = fmaf(mx, 6.9053396600248786e-4f, 6.9053396600248786e-4f);  // 2^-10.5 (1 + max)
    uint16_t* out = reinterpret_cast<uint16_t*>(p.bits + 2 * frame * (size_t)N);
    if (mn > uu * uu) {
#pragma unroll
      for (int c = 0; c < 32; ++c) out[t + 32 * c] = (uint16_t)x2_sign_pair(y[c], mask);
    } else {
      // rare: near an axis, NaN or inf somewhere in this thread's bins -> per-symbol test, exact path where it fails
#pragma unroll 1
      for (int c = 0; c < 32; ++c) {
        float2 s = y[0];
#pragma unroll
        for (int k = 1; k < 32; ++k) s = (c == k) ? y[k] : s;   // select without dynamic register indexing
        out[t + 32 * c] = (uint16_t)(qpsk_fast_ok(s) ? x2_sign_pair(s, mask) : qpsk_pair_from_index(demod_qpsk_exact_slow(s), hi_shift));
      }
    }
    x2_syncwarp();                                         // the exchange buffer is rewritten by the next frame
  }
}

}  // namespace ae
