// elementwise.cu — HBM-bound element-wise kernels of the cf32 path (sm_100a):
//   K1  vecops_fused      src/vecops.rs:94-182  (op-tape interpreter, one pass for a whole chain)
//   K5  downsample        src/sampling.rs:28-62
//   K6  interpolate       src/sampling.rs:7-24
//   K7  modulate          src/modulation.rs:115-131
//   K8  demod_hard        src/modulation.rs:33-56, :133-144
//   K9  awgn_fill/apply   src/noise.rs:39-66    (Philox4x32-10 + Box-Muller, counter = sample index)
//   K10 modem_fused       examples/modem.rs:15-32
//   K11 expand / mseq     src/sequence.rs:18-53
//   K13 bit-error / EVM partial sums
// All float arithmetic that the reference's tests pin exactly uses the *_exact helpers
// (no FMA contraction, IEEE division, denormals kept).
#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace ae {

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// =================================================================================================
// K1 fused VecOps
// =================================================================================================
template <int W> struct VecIO;
template <> struct VecIO<1> {
  static __device__ __forceinline__ void load(float2 (&d)[1], const float2* p) { d[0] = ld_stream(p); }
  static __device__ __forceinline__ void store(float2* p, const float2 (&d)[1]) { st_stream(p, d[0]); }
};
template <> struct VecIO<2> {
  static __device__ __forceinline__ void load(float2 (&d)[2], const float2* p) {
    const float4 v = ld_stream(reinterpret_cast<const float4*>(p));
    d[0] = make_float2(v.x, v.y); d[1] = make_float2(v.z, v.w);
  }
  static __device__ __forceinline__ void store(float2* p, const float2 (&d)[2]) {
    st_stream(reinterpret_cast<float4*>(p), make_float4(d[0].x, d[0].y, d[1].x, d[1].y));
  }
};

template <> struct VecIO<4> {   // 256-bit accesses (sm_100 LDG.256 / STG.256), 32-byte aligned
  static __device__ __forceinline__ void load(float2 (&d)[4], const float2* p) {
    float4 a, b;
    ld_stream_256(p, a, b);
    d[0] = make_float2(a.x, a.y); d[1] = make_float2(a.z, a.w); d[2] = make_float2(b.x, b.y); d[3] = make_float2(b.z, b.w);
  }
  static __device__ __forceinline__ void store(float2* p, const float2 (&d)[4]) {
    st_stream_256(p, make_float4(d[0].x, d[0].y, d[1].x, d[1].y), make_float4(d[2].x, d[2].y, d[3].x, d[3].y));
  }
};

template <int W>
__device__ __forceinline__ void tape_apply(float2 (&x)[W], const TapeEntry& e, size_t idx) {
  float2 o[W];
  switch (e.op) {
    case OP_SCALE:
#pragma unroll
      for (int w = 0; w < W; ++w) x[w] = cx_scale_exact(x[w], e.s);
      break;
    case OP_CONJ:
#pragma unroll
      for (int w = 0; w < W; ++w) x[w].y = -x[w].y;
      break;
    case OP_ZERO:
#pragma unroll
      for (int w = 0; w < W; ++w) x[w] = make_float2(0.0f, 0.0f);
      break;
    case OP_MIRROR: break;  // handled by the caller (register swap)
    default:
      VecIO<W>::load(o, e.operand + idx);
#pragma unroll
      for (int w = 0; w < W; ++w) {
        if (e.op == OP_MUL) x[w] = cx_mul_exact(x[w], o[w]);
        else if (e.op == OP_DIV) x[w] = cx_div_exact(x[w], o[w]);
        else if (e.op == OP_ADD) x[w] = cx_add_exact(x[w], o[w]);
        else if (e.op == OP_SUB) x[w] = cx_sub_exact(x[w], o[w]);
        else x[w] = o[w];  // OP_CLONE
      }
  }
}

constexpr int kVecThreads = 256;
__host__ __device__ constexpr int vec_unroll(int w) { return w == 4 ? 2 : 4; }   // 64 bytes per operand in flight per thread

// no mirror in the tape: W consecutive samples per pack, kVecUnroll packs per thread
template <int W>
__global__ void __launch_bounds__(kVecThreads) vecops_kernel(float2* __restrict__ v, size_t n, const __grid_constant__ TapeParams p) {
  constexpr int kVecUnroll = vec_unroll(W);
  const size_t packs = n / W;
  const size_t base = (size_t)blockIdx.x * (kVecThreads * kVecUnroll) + threadIdx.x;
  float2 x[kVecUnroll][W];
#pragma unroll
  for (int u = 0; u < kVecUnroll; ++u) {
    const size_t i = base + (size_t)u * kVecThreads;
    if (i < packs && p.load_self) VecIO<W>::load(x[u], v + i * W);
    else {
#pragma unroll
      for (int w = 0; w < W; ++w) x[u][w] = make_float2(0.0f, 0.0f);
    }
  }
  for (int k = 0; k < p.n_ops; ++k) {
#pragma unroll
    for (int u = 0; u < kVecUnroll; ++u) {
      const size_t i = base + (size_t)u * kVecThreads;
      if (i < packs) tape_apply<W>(x[u], p.e[k], i * W);
    }
  }
#pragma unroll
  for (int u = 0; u < kVecUnroll; ++u) {
    const size_t i = base + (size_t)u * kVecThreads;
    if (i < packs) VecIO<W>::store(v + i * W, x[u]);
  }
  if (W > 1 && blockIdx.x == 0 && threadIdx.x < n % W) {  // tail elements that do not fill a pack
    float2 t[1];
    const size_t i = packs * W + threadIdx.x;
    t[0] = p.load_self ? v[i] : make_float2(0.0f, 0.0f);
    for (int k = 0; k < p.n_ops; ++k) tape_apply<1>(t, p.e[k], i);
    v[i] = t[0];
  }
}

// tape contains vec_mirror: one thread owns the pair (p, p+mid) so the swap is a register
// rename and operands are always read at the positions the data currently occupies.
template <int W>
__global__ void __launch_bounds__(kVecThreads) vecops_mirror_kernel(float2* __restrict__ v, size_t n, const __grid_constant__ TapeParams p) {
  const size_t mid = n / 2;
  const size_t packs = mid / W;  // host guarantees mid % W == 0
  const size_t i = (size_t)blockIdx.x * kVecThreads + threadIdx.x;
  if (i < packs) {
    float2 lo[W], hi[W];
    const size_t a = i * W, b = mid + i * W;
    if (p.load_self) { VecIO<W>::load(lo, v + a); VecIO<W>::load(hi, v + b); }
    else {
#pragma unroll
      for (int w = 0; w < W; ++w) { lo[w] = make_float2(0.0f, 0.0f); hi[w] = lo[w]; }
    }
    for (int k = 0; k < p.n_ops; ++k) {
      if (p.e[k].op == OP_MIRROR) {
#pragma unroll
        for (int w = 0; w < W; ++w) { const float2 t = lo[w]; lo[w] = hi[w]; hi[w] = t; }
      } else {
        tape_apply<W>(lo, p.e[k], a);
        tape_apply<W>(hi, p.e[k], b);
      }
    }
    VecIO<W>::store(v + a, lo);
    VecIO<W>::store(v + b, hi);
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd length: last element never moves (:157-161)
    float2 t[1];
    const size_t j = n - 1;
    t[0] = p.load_self ? v[j] : make_float2(0.0f, 0.0f);
    for (int k = 0; k < p.n_ops; ++k) tape_apply<1>(t, p.e[k], j);
    v[j] = t[0];
  }
}

void launch_vecops(float2* v, size_t n, const TapeParams& p, bool has_mirror, int, cudaStream_t st) {
  if (n == 0) return;
  bool aligned = ((uintptr_t)v % 16) == 0, aligned32 = ((uintptr_t)v % 32) == 0;
  for (int k = 0; k < p.n_ops; ++k)
    if (p.e[k].operand) {
      if (((uintptr_t)p.e[k].operand % 16) != 0) aligned = false;
      if (((uintptr_t)p.e[k].operand % 32) != 0) aligned32 = false;
    }
  if (!has_mirror) {
    if (aligned32 && n >= 4) {
      vecops_kernel<4><<<cdiv(n / 4, kVecThreads * vec_unroll(4)), kVecThreads, 0, st>>>(v, n, p);
    } else if (aligned && n >= 2) {
      vecops_kernel<2><<<cdiv(n / 2, kVecThreads * vec_unroll(2)), kVecThreads, 0, st>>>(v, n, p);
    } else {
      vecops_kernel<1><<<cdiv(n, kVecThreads * vec_unroll(1)), kVecThreads, 0, st>>>(v, n, p);
    }
  } else {
    const size_t mid = n / 2;
    if (aligned32 && (mid % 4) == 0 && mid >= 4) {
      vecops_mirror_kernel<4><<<cdiv(mid / 4, kVecThreads), kVecThreads, 0, st>>>(v, n, p);
    } else if (aligned && (mid % 2) == 0 && mid >= 2) {
      vecops_mirror_kernel<2><<<cdiv(mid / 2, kVecThreads), kVecThreads, 0, st>>>(v, n, p);
    } else {
      vecops_mirror_kernel<1><<<cdiv(mid ? mid : 1, kVecThreads), kVecThreads, 0, st>>>(v, n, p);
    }
  }
}

// =================================================================================================
// K5 downsample (src/sampling.rs:38-41): dst[i] = src[i*dec]
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) downsample_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t n_dst, size_t dec) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n_dst) dst[i] = src[i * dec];
}
// dec == 2 / 4 with 16-byte aligned cf32: each thread reads whole 16-byte vectors of the source
// it needs and writes one float4 (two outputs) so every store instruction is a full 512-byte row.
template <int DEC>
__global__ void __launch_bounds__(256) downsample_cf32_vec_kernel(const float2* __restrict__ src, float2* __restrict__ dst, size_t n_pairs) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < n_pairs) {
    const float2 a = __ldcs(src + (2 * i) * DEC);
    const float2 b = __ldcs(src + (2 * i + 1) * DEC);
    __stcs(reinterpret_cast<float4*>(dst) + i, make_float4(a.x, a.y, b.x, b.y));
  }
}
void launch_downsample_cf32(const float2* src, float2* dst, size_t n_dst, size_t dec, cudaStream_t st) {
  if (n_dst == 0) return;
  const bool al = ((uintptr_t)dst % 16) == 0;
  const size_t pairs = n_dst / 2;
  if (al && pairs && (dec == 4 || dec == 2)) {
    if (dec == 4) downsample_cf32_vec_kernel<4><<<cdiv(pairs, 256), 256, 0, st>>>(src, dst, pairs);
    else downsample_cf32_vec_kernel<2><<<cdiv(pairs, 256), 256, 0, st>>>(src, dst, pairs);
    if (n_dst & 1) downsample_kernel<float2><<<1, 256, 0, st>>>(src + (n_dst - 1) * dec, dst + (n_dst - 1), 1, dec);
  } else {
    downsample_kernel<float2><<<cdiv(n_dst, 256), 256, 0, st>>>(src, dst, n_dst, dec);
  }
}
void launch_downsample_u8(const uint8_t* src, uint8_t* dst, size_t n_dst, size_t dec, cudaStream_t st) {
  if (n_dst == 0) return;
  downsample_kernel<uint8_t><<<cdiv(n_dst, 256), 256, 0, st>>>(src, dst, n_dst, dec);
}

// =================================================================================================
// K6 interpolate (src/sampling.rs:7-24).  One thread per input window; K1 = n_between+1 outputs.
//   rate = ((x2.re-x1.re)/K1, (x2.im-x1.im)/K1);  out = (x1.re + i*rate.0, x1.{re|im} + i*rate.1)
// =================================================================================================
template <int K1T>  // 0 = runtime K1
__global__ void __launch_bounds__(256) interpolate_kernel(const float2* __restrict__ src, size_t n_src, float2* __restrict__ dst,
                                                          int k1_rt, int compat, int vec_ok) {
  const int K1 = K1T ? K1T : k1_rt;
  const size_t w = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (w + 1 < n_src) {
    const float2 x1 = __ldg(src + w), x2 = __ldg(src + w + 1);
    const float div = (float)K1;
    const float r0 = __fdiv_rn(__fsub_rn(x2.x, x1.x), div);
    const float r1 = __fdiv_rn(__fsub_rn(x2.y, x1.y), div);
    const float imb = compat == AE_COMPAT_REFERENCE ? x1.x : x1.y;  // :19 uses x1.re
    float2* o = dst + w * (size_t)K1;
    if (K1T == 4 && vec_ok == 2) {
      // 32 bytes per window: one 256-bit store (sm_100 STG.256), so a warp store covers 1 KiB contiguously
      float v[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float f = (float)i;
        v[2 * i] = __fadd_rn(x1.x, __fmul_rn(f, r0));
        v[2 * i + 1] = __fadd_rn(imb, __fmul_rn(f, r1));
      }
      st_stream_256(o, make_float4(v[0], v[1], v[2], v[3]), make_float4(v[4], v[5], v[6], v[7]));
    } else if (K1T > 0 && (K1T % 2) == 0 && vec_ok) {
#pragma unroll
      for (int i = 0; i < K1T; i += 2) {
        const float f0 = (float)i, f1 = (float)(i + 1);
        const float4 v = make_float4(__fadd_rn(x1.x, __fmul_rn(f0, r0)), __fadd_rn(imb, __fmul_rn(f0, r1)),
                                     __fadd_rn(x1.x, __fmul_rn(f1, r0)), __fadd_rn(imb, __fmul_rn(f1, r1)));
        __stcs(reinterpret_cast<float4*>(o + i), v);
      }
    } else {
      for (int i = 0; i < K1; ++i) {
        const float f = (float)i;
        __stcs(o + i, make_float2(__fadd_rn(x1.x, __fmul_rn(f, r0)), __fadd_rn(imb, __fmul_rn(f, r1))));
      }
    }
  } else if (w + 1 == n_src) {
    dst[w * (size_t)K1] = src[w];  // dst.push(*src.last().unwrap()) :23
  }
}
void launch_interpolate(const float2* src, size_t n_src, float2* dst, size_t n_between, int compat, cudaStream_t st) {
  if (n_src == 0) return;
  const int k1 = (int)(n_between + 1);
  const int vec_ok = ((uintptr_t)dst % 32) == 0 ? 2 : (((uintptr_t)dst % 16) == 0 ? 1 : 0);   // 2: 256-bit stores allowed
  const unsigned g = cdiv(n_src, 256);
  if (k1 == 4) interpolate_kernel<4><<<g, 256, 0, st>>>(src, n_src, dst, k1, compat, vec_ok);
  else if (k1 == 2) interpolate_kernel<2><<<g, 256, 0, st>>>(src, n_src, dst, k1, compat, vec_ok);
  else if (k1 == 8) interpolate_kernel<8><<<g, 256, 0, st>>>(src, n_src, dst, k1, compat, vec_ok);
  else interpolate_kernel<0><<<g, 256, 0, st>>>(src, n_src, dst, k1, compat, 0);
}

// =================================================================================================
// K7 modulate (src/modulation.rs:115-121): chunks(BPS) -> index -> LUT
// =================================================================================================
template <int BPS>
__device__ __forceinline__ float2 mod_symbol(const ModTable& tab, const uint8_t* b, int* errflag) {
  unsigned idx;
  if (BPS == 1) idx = b[0];                                                 // :10
  else idx = (uint8_t)((uint8_t)(b[1] << 1) + b[0]);                        // :24 (u8 arithmetic)
  if (idx >= (unsigned)tab.len) { atomicOr(errflag, DEVERR_MOD_INDEX); return make_float2(0.0f, 0.0f); }
  return tab.t[idx];
}
template <int BPS>
__global__ void __launch_bounds__(256) modulate_kernel(const __grid_constant__ ModTable tab, const uint8_t* __restrict__ bits,
                                                       float2* __restrict__ out, size_t n_out, int vec_ok, int* errflag) {
  const size_t g = (size_t)blockIdx.x * 256 + threadIdx.x;  // group of 4 symbols
  const size_t s0 = g * 4;
  if (s0 >= n_out) return;
  if (vec_ok && s0 + 4 <= n_out) {
    uint8_t b[4 * BPS];
    if (BPS == 2) *reinterpret_cast<uint2*>(b) = __ldcs(reinterpret_cast<const uint2*>(bits + s0 * 2));
    else *reinterpret_cast<uint32_t*>(b) = __ldcs(reinterpret_cast<const uint32_t*>(bits + s0));
    float2 s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = mod_symbol<BPS>(tab, b + i * BPS, errflag);
    float4* o = reinterpret_cast<float4*>(out + s0);
    if (vec_ok == 2) {
      st_stream_256(o, make_float4(s[0].x, s[0].y, s[1].x, s[1].y), make_float4(s[2].x, s[2].y, s[3].x, s[3].y));
    } else {
      __stcs(o, make_float4(s[0].x, s[0].y, s[1].x, s[1].y));
      __stcs(o + 1, make_float4(s[2].x, s[2].y, s[3].x, s[3].y));
    }
  } else {
    for (size_t s = s0; s < n_out && s < s0 + 4; ++s) out[s] = mod_symbol<BPS>(tab, bits + s * BPS, errflag);
  }
}
void launch_modulate(const ModTable& tab, const uint8_t* bits, size_t, float2* out, size_t n_out, int* errflag, cudaStream_t st) {
  if (n_out == 0) return;
  const int vec_ok = (((uintptr_t)bits % 8) == 0 && ((uintptr_t)out % 16) == 0) ? (((uintptr_t)out % 32) == 0 ? 2 : 1) : 0;   // 2: 256-bit stores
  const unsigned g = cdiv(cdiv(n_out, 4), 256);
  if (tab.len == 2) modulate_kernel<1><<<g, 256, 0, st>>>(tab, bits, out, n_out, vec_ok, errflag);
  else modulate_kernel<2><<<g, 256, 0, st>>>(tab, bits, out, n_out, vec_ok, errflag);
}

// =================================================================================================
// K8 hard demod.  BPSK: generic default impl (:133-144); QPSK: the override (:33-56) which emits
// idx&1 then idx&2 in {0,2}; compat=corrected emits (idx>>1)&1.
// =================================================================================================
template <int M>
__device__ __forceinline__ void demod_emit(float2 s, const float2 (&tab)[M], int compat, uint8_t* o, bool generic = false) {
  const unsigned idx = (M == 4 && generic) ? demod_qpsk_generic(s) : demod_index<M>(s, tab);
  o[0] = (uint8_t)(idx & 1u);
  if (M == 4) o[1] = compat == AE_COMPAT_REFERENCE ? (uint8_t)(idx & 2u) : (uint8_t)((idx >> 1) & 1u);
}
template <int M>
__global__ void __launch_bounds__(256) demod_kernel(const __grid_constant__ ModTable tabp, const float2* __restrict__ sym, size_t n,
                                                    uint8_t* __restrict__ bits, int compat, int vec_ok) {
  constexpr int BPS = M == 2 ? 1 : 2;
  float2 tab[M];
#pragma unroll
  for (int c = 0; c < M; ++c) tab[c] = tabp.t[c];
  const bool generic = tabp.generic_qpsk != 0;
  const size_t s0 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (s0 >= n) return;
  if (vec_ok && s0 + 4 <= n) {
    float4 a, b;
    if (vec_ok == 2) {
      ld_stream_256(sym + s0, a, b);                       // four symbols in one 256-bit load
    } else {
      a = __ldcs(reinterpret_cast<const float4*>(sym + s0));
      b = __ldcs(reinterpret_cast<const float4*>(sym + s0) + 1);
    }
    uint8_t o[4 * BPS];
    demod_emit<M>(make_float2(a.x, a.y), tab, compat, o, generic);
    demod_emit<M>(make_float2(a.z, a.w), tab, compat, o + BPS, generic);
    demod_emit<M>(make_float2(b.x, b.y), tab, compat, o + 2 * BPS, generic);
    demod_emit<M>(make_float2(b.z, b.w), tab, compat, o + 3 * BPS, generic);
    if (BPS == 2) __stcs(reinterpret_cast<uint2*>(bits + s0 * 2), *reinterpret_cast<uint2*>(o));
    else __stcs(reinterpret_cast<uint32_t*>(bits + s0), *reinterpret_cast<uint32_t*>(o));
  } else {
    for (size_t s = s0; s < n && s < s0 + 4; ++s) demod_emit<M>(sym[s], tab, compat, bits + s * BPS, generic);
  }
}
void launch_demod(const ModTable& tab, const float2* sym, size_t n, uint8_t* bits, int compat, cudaStream_t st) {
  if (n == 0) return;
  const int vec_ok = (((uintptr_t)sym % 16) == 0 && ((uintptr_t)bits % 8) == 0) ? (((uintptr_t)sym % 32) == 0 ? 2 : 1) : 0;   // 2: 256-bit loads
  const unsigned g = cdiv(cdiv(n, 4), 256);
  if (tab.len == 2) demod_kernel<2><<<g, 256, 0, st>>>(tab, sym, n, bits, compat, vec_ok);
  else demod_kernel<4><<<g, 256, 0, st>>>(tab, sym, n, bits, compat, vec_ok);
}

// =================================================================================================
// K9 AWGN (src/noise.rs:39-66).  Thread = one Philox block = the sample pair (2P, 2P+1) of the
// stream; results depend only on (seed, stream, global sample index), never on the grid shape.
// =================================================================================================
template <bool APPLY>
__global__ void __launch_bounds__(256) awgn_kernel(float2* __restrict__ buf, size_t n, float scale, int twice,
                                                   const __grid_constant__ PhiloxKeys keys, uint64_t stream, uint64_t offset, int vec_ok) {
  // thread = two consecutive Philox blocks = the four samples 4Q .. 4Q+3 of the stream (32 bytes)
  const uint64_t q0 = offset >> 2;
  const uint64_t Q = q0 + (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const uint64_t g0 = 4 * Q;
  if (g0 >= offset + n) return;
  float2 z[4];
  awgn_unit_pair(keys, stream, 2 * Q, z[0], z[1]);
  awgn_unit_pair(keys, stream, 2 * Q + 1, z[2], z[3]);
  // next(): (N(0,1) as f32) * scale (:41-42); apply(): ... .scale(sc) once more (:58)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    z[i] = cx_scale_exact(z[i], scale);
    if (twice) z[i] = cx_scale_exact(z[i], scale);
  }
  if (vec_ok && g0 >= offset && g0 + 4 <= offset + n) {
    float4* q = reinterpret_cast<float4*>(buf + (g0 - offset));
    float4 a = make_float4(z[0].x, z[0].y, z[1].x, z[1].y), b = make_float4(z[2].x, z[2].y, z[3].x, z[3].y);
    if (APPLY) {
      float4 s0, s1;
      ld_stream_256(q, s0, s1);
      a = make_float4(__fadd_rn(s0.x, a.x), __fadd_rn(s0.y, a.y), __fadd_rn(s0.z, a.z), __fadd_rn(s0.w, a.w));
      b = make_float4(__fadd_rn(s1.x, b.x), __fadd_rn(s1.y, b.y), __fadd_rn(s1.z, b.z), __fadd_rn(s1.w, b.w));
    }
    st_stream_256(q, a, b);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint64_t g = g0 + i;
      if (g >= offset && g < offset + n) { float2* q = buf + (g - offset); *q = APPLY ? cx_add_exact(*q, z[i]) : z[i]; }
    }
  }
}
static void awgn_launch(bool apply, float2* buf, size_t n, float scale, int twice, uint64_t seed, uint64_t stream,
                        uint64_t offset, cudaStream_t st) {
  if (n == 0) return;
  const uint64_t quads = ((offset + n + 3) >> 2) - (offset >> 2);
  const int vec_ok = ((offset & 3) == 0) && ((uintptr_t)buf % 32) == 0;
  const unsigned g = cdiv(quads, 256);
  const PhiloxKeys keys = make_philox_keys(seed);
  if (apply) awgn_kernel<true><<<g, 256, 0, st>>>(buf, n, scale, twice, keys, stream, offset, vec_ok);
  else awgn_kernel<false><<<g, 256, 0, st>>>(buf, n, scale, twice, keys, stream, offset, vec_ok);
}
void launch_awgn_fill(float2* dst, size_t n, float scale, uint64_t seed, uint64_t stream, uint64_t offset, cudaStream_t st) {
  awgn_launch(false, dst, n, scale, 0, seed, stream, offset, st);
}
void launch_awgn_apply(float2* sig, size_t n, float scale, int twice, uint64_t seed, uint64_t stream, uint64_t offset,
                       cudaStream_t st) {
  awgn_launch(true, sig, n, scale, twice, seed, stream, offset, st);
}

// =================================================================================================
// warp / block reductions for the statistics
// =================================================================================================
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {   // selector nibble bit 3: replicate the byte's sign
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ float fmax3_nan(float a, float b, float c) {   // FMNMX3, NaN-propagating
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// =================================================================================================
// K10 modem_fused (examples/modem.rs:15-32): modulate -> Awgn::apply -> demod_naive (+ bit errors)
// in registers.  4 B/symbol of HBM traffic for QPSK: 2 B bits in + 2 B bits out.
// Thread = 4 symbols (two Philox blocks).
// =================================================================================================
template <int M, bool TWICE>
__global__ void __launch_bounds__(256) modem_kernel(const __grid_constant__ ModTable tabp, const uint8_t* __restrict__ bin, size_t nsym,
                                                    uint8_t* __restrict__ bout, float scale, const __grid_constant__ PhiloxKeys keys,
                                                    uint64_t stream, uint64_t offset, int compat, ae_stats* stats, int* errflag, int vec_ok) {
  constexpr int BPS = M == 2 ? 1 : 2;
  float2 tab[M];
#pragma unroll
  for (int c = 0; c < M; ++c) tab[c] = tabp.t[c];
  const bool generic = tabp.generic_qpsk != 0;
  unsigned long long errs = 0;
  int neg_errs = 0;       // fast path: minus the bit errors of this thread (<= 8 per iteration, far from overflow)
  // pairs are aligned to the GLOBAL sample index so the stream does not depend on `offset` parity
  const uint64_t p0 = offset >> 1;
  const uint64_t npairs = ((offset + nsym + 1) >> 1) - p0;
  for (uint64_t q = ((uint64_t)blockIdx.x * 256 + threadIdx.x) * 2; q < npairs; q += (uint64_t)gridDim.x * 512) {
    alignas(8) uint8_t bi[4 * BPS];
    alignas(8) uint8_t bo[4 * BPS];
    const uint64_t g0 = 2 * (p0 + q);           // global index of the first of 4 samples
    const long long l0 = (long long)(g0 - offset);  // local symbol index (may be -1)
    const bool full = vec_ok && l0 >= 0 && (size_t)(l0 + 4) <= nsym;
    if (full) {
      if (BPS == 2) *reinterpret_cast<uint2*>(bi) = __ldcs(reinterpret_cast<const uint2*>(bin + l0 * 2));
      else *reinterpret_cast<uint32_t*>(bi) = __ldcs(reinterpret_cast<const uint32_t*>(bin + l0));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long l = l0 + i;
        const bool ok = l >= 0 && (size_t)l < nsym;
#pragma unroll
        for (int b = 0; b < BPS; ++b) bi[i * BPS + b] = ok ? bin[l * BPS + b] : 0;
      }
    }
    float2 z[4];
    awgn_unit_pair(keys, stream, p0 + q, z[0], z[1]);
    awgn_unit_pair(keys, stream, p0 + q + 1, z[2], z[3]);
    // CHECKED = false: all four symbols exist (the common case) - no per-symbol bounds tests, bit errors
    // counted on the packed bytes afterwards
    auto body = [&](auto checked_tag) {
      constexpr bool CHECKED = decltype(checked_tag)::value;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long l = l0 + i;
        const bool ok = !CHECKED || (l >= 0 && (size_t)l < nsym);
        unsigned idx;
        if (BPS == 1) idx = bi[i];
        else idx = (uint8_t)((uint8_t)(bi[2 * i + 1] << 1) + bi[2 * i]);
        float2 s = make_float2(0.0f, 0.0f);
        if (idx < (unsigned)M) s = tabp.t[idx];   // indexed read of the constant bank (tab[] is a register copy for demod)
        else if (ok) atomicOr(errflag, DEVERR_MOD_INDEX);
        float2 nz = cx_scale_exact(z[i], scale);
        if (TWICE) nz = cx_scale_exact(nz, scale);
        s = cx_add_exact(s, nz);
        demod_emit<M>(s, tab, compat, bo + i * BPS, generic);
        if (CHECKED && ok) {
#pragma unroll
          for (int b = 0; b < BPS; ++b) errs += ((bi[i * BPS + b] != 0) != (bo[i * BPS + b] != 0));
        }
      }
    };
    // Fast path for the crate's own QPSK table (+-1, +-1) with every bit byte 0 or 1: the constellation point is
    // sigma = 1 - 2 bit per component and fl(sigma + n) = sigma * fl(1 + sigma n) exactly (IEEE rounding is symmetric),
    // so one FADD on the sign-flipped noise gives |component| and, as its sign, "this bit was received wrong".
    // The reference's nearest-point search agrees with the sign test unless a component is within the bound of
    // demod_qpsk_generic() of an axis, NaN or inf: one test per thread on the minimum and NaN-propagating maximum of
    // its eight components, the general code below otherwise.
    bool fast_done = false;
    if (M == 4 && full && generic) {
      const uint2 w = *reinterpret_cast<const uint2*>(bi);
      if (((w.x | w.y) & 0xfefefefeu) == 0u) {
        const uint32_t sx = w.x << 7, sy = w.y << 7;       // bit 7 of every byte = the bit
        float r[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 nz = cx_scale_exact(z[i], scale);
          if (TWICE) nz = cx_scale_exact(nz, scale);
          const uint32_t src = i < 2 ? sx : sy;
          const uint32_t mre = prmt_b32(src, 0u, 0x8888u + 0x1111u * (2 * (i & 1)));       // all ones when the bit is set
          const uint32_t mim = prmt_b32(src, 0u, 0x8888u + 0x1111u * (2 * (i & 1) + 1));
          r[2 * i] = __fadd_rn(1.0f, __uint_as_float(__float_as_uint(nz.x) ^ (mre & 0x80000000u)));
          r[2 * i + 1] = __fadd_rn(1.0f, __uint_as_float(__float_as_uint(nz.y) ^ (mim & 0x80000000u)));
        }
        float mx = fmax3_nan(fabsf(r[0]), fabsf(r[1]), fabsf(r[2])), mn = fmin3(fabsf(r[0]), fabsf(r[1]), fabsf(r[2]));
        mx = fmax3_nan(mx, fabsf(r[3]), fabsf(r[4])); mn = fmin3(mn, fabsf(r[3]), fabsf(r[4]));
        mx = fmax3_nan(mx, fabsf(r[5]), fabsf(r[6])); mn = fmin3(mn, fabsf(r[5]), fabsf(r[6]));
        mx = fmax3_nan(mx, fabsf(r[7]), fabsf(r[7])); mn = fmin3(mn, fabsf(r[7]), fabsf(r[7]));
        const float uu = fmaf(mx, 6.9053396600248786e-4f, 6.9053396600248786e-4f);          // 2^-10.5 (1 + max)
        if (mn > uu * uu) {
          // bytes of e: 0xff where the received bit differs from the sent one
          const uint32_t e0 = prmt_b32(prmt_b32(__float_as_uint(r[0]), __float_as_uint(r[1]), 0x00FBu),
                                       prmt_b32(__float_as_uint(r[2]), __float_as_uint(r[3]), 0x00FBu), 0x5410u);
          const uint32_t e1 = prmt_b32(prmt_b32(__float_as_uint(r[4]), __float_as_uint(r[5]), 0x00FBu),
                                       prmt_b32(__float_as_uint(r[6]), __float_as_uint(r[7]), 0x00FBu), 0x5410u);
          uint32_t o0 = w.x ^ (e0 & 0x01010101u), o1 = w.y ^ (e1 & 0x01010101u);
          if (compat == AE_COMPAT_REFERENCE) { o0 += o0 & 0x01000100u; o1 += o1 & 0x01000100u; }   // second byte = idx & 2
          *reinterpret_cast<uint2*>(bo) = make_uint2(o0, o1);
          // bytes of e are 0 or -1: a signed byte dot product with ones subtracts the count (IDP.4A; POPC costs 8 dispatch cycles)
          neg_errs = __dp4a((int)e1, 0x01010101, __dp4a((int)e0, 0x01010101, neg_errs));
          fast_done = true;
        }
      }
    }
    if (fast_done) {
    } else if (full) {
      body(std::false_type{});
      // a bit is in error when exactly one of (sent, received) is non-zero: compare the non-zero masks
      if (BPS == 2) {
        const uint2 a = *reinterpret_cast<const uint2*>(bi), b = *reinterpret_cast<const uint2*>(bo);
        errs += (__popc(__vcmpne4(a.x, 0u) ^ __vcmpne4(b.x, 0u)) + __popc(__vcmpne4(a.y, 0u) ^ __vcmpne4(b.y, 0u))) >> 3;
      } else {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(bi), b = *reinterpret_cast<const uint32_t*>(bo);
        errs += __popc(__vcmpne4(a, 0u) ^ __vcmpne4(b, 0u)) >> 3;
      }
    } else {
      body(std::true_type{});
    }
    if (full) {
      if (BPS == 2) __stcs(reinterpret_cast<uint2*>(bout + l0 * 2), *reinterpret_cast<uint2*>(bo));
      else __stcs(reinterpret_cast<uint32_t*>(bout + l0), *reinterpret_cast<uint32_t*>(bo));
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long l = l0 + i;
        if (l >= 0 && (size_t)l < nsym) {
#pragma unroll
          for (int b = 0; b < BPS; ++b) bout[l * BPS + b] = bo[i * BPS + b];
        }
      }
    }
  }
  if (stats) {
    errs = warp_sum_u64(errs + (unsigned long long)(unsigned)(-neg_errs));
    if ((threadIdx.x & 31) == 0 && errs) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->bit_errors), errs);
    if (blockIdx.x == 0 && threadIdx.x == 0)
      atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_bits), (unsigned long long)nsym * BPS);
  }
}
void launch_modem_fused(const ModTable& tab, const uint8_t* bits_in, size_t nbits, uint8_t* bits_out, float scale, int twice,
                        uint64_t seed, uint64_t stream, uint64_t offset, int compat, ae_stats* stats, int* errflag,
                        int sm_count, cudaStream_t st) {
  const int bps = tab.len == 2 ? 1 : 2;
  const size_t nsym = nbits / bps;
  if (nsym == 0) return;
  const int vec_ok = ((uintptr_t)bits_in % 8) == 0 && ((uintptr_t)bits_out % 8) == 0 && (offset % 2) == 0;
  const uint64_t npairs = ((offset + nsym + 1) >> 1) - (offset >> 1);
  unsigned g = cdiv(cdiv(npairs, 2), 256);
  const unsigned cap = (unsigned)sm_count * 16;
  if (g > cap) g = cap;
  const PhiloxKeys keys = make_philox_keys(seed);
  auto go = [&](auto kern) { kern<<<g, 256, 0, st>>>(tab, bits_in, nsym, bits_out, scale, keys, stream, offset, compat, stats, errflag, vec_ok); };
  if (tab.len == 2) { if (twice) go(modem_kernel<2, true>); else go(modem_kernel<2, false>); }
  else { if (twice) go(modem_kernel<4, true>); else go(modem_kernel<4, false>); }
}

// =================================================================================================
// K11 sequence
// =================================================================================================
__global__ void expand_kernel(uint64_t seed, size_t len, uint8_t* out) {  // src/sequence.rs:18-21
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i < len) out[i] = (uint8_t)((seed >> i) & 1u);
}
void launch_expand(uint64_t seed, size_t len, uint8_t* out, cudaStream_t st) {
  if (len) expand_kernel<<<cdiv(len, 256), 256, 0, st>>>(seed, len, out);
}

// GF(2)[z] arithmetic modulo p(z) = z^deg + poly_low(z), deg <= 64
__device__ __forceinline__ uint64_t gf2_mulmod(uint64_t a, uint64_t b, uint64_t poly_low, int deg) {
  // shift-and-add with reduction folded into every doubling of `a`
  uint64_t r = 0;
  const uint64_t top = 1ull << (deg - 1);
  for (int i = 0; i < deg; ++i) {
    if ((b >> i) & 1ull) r ^= a;
    const bool carry = (a & top) != 0;
    a = (deg == 64) ? (a << 1) : ((a << 1) & ((1ull << deg) - 1ull));
    if (carry) a ^= poly_low;
  }
  return r;
}
constexpr int kMseqBitsPerThread = 512;
constexpr int kMseqThreads = 128;
constexpr int kMseqPitch = kMseqThreads + 2;   // half-words per 16-bit group row in shared memory
struct MseqTables {
  uint64_t pow2[64];              // z^(2^i) mod p
  uint64_t thread[kMseqThreads];  // z^(kMseqBitsPerThread * t) mod p
};
// 4 bits -> 4 bytes of 0/1 (little endian)
__device__ __forceinline__ uint32_t spread4(uint32_t b) { return ((b & 0xfu) * 0x00204081u) & 0x01010101u; }

// out[i] = x[i], i < len, where x obeys x[n] = parity(window & poly_low) with window bit j = x[n-deg+j]
// and x[0..deg) = state bits.  Jump-ahead: z^n mod p(z) = sum c_i z^i  =>  x[n] = parity(c & state).
// Block start: product of the host-tabulated z^(2^i) over the set bits of the offset (thread 0);
// thread start: one more multiplication with the tabulated z^(512 t); then 512 bits per thread from a
// 64-bit sliding window, packed 16 to a half-word in shared memory and expanded by the whole CTA.
__device__ __forceinline__ void mseq_fill_bits(uint64_t state, uint64_t poly_low, int deg, size_t len, const MseqTables& tab,
                                               uint16_t* s_bits) {
  __shared__ uint64_t s_block;
  const size_t block_start = (size_t)blockIdx.x * kMseqThreads * kMseqBitsPerThread;
  if (threadIdx.x == 0) {
    uint64_t r = 1ull;
    size_t e = block_start;
    for (int i = 0; e; ++i, e >>= 1)
      if (e & 1) r = gf2_mulmod(r, tab.pow2[i], poly_low, deg);
    s_block = r;
  }
  __syncthreads();
  const size_t n0 = block_start + (size_t)threadIdx.x * kMseqBitsPerThread;
  if (n0 >= len) return;
  uint64_t r = gf2_mulmod(tab.thread[threadIdx.x], s_block, poly_low, deg);
  // first deg bits of this thread's run via successive multiplication by z
  const uint64_t top = 1ull << (deg - 1);
  uint64_t window = 0;
  for (int j = 0; j < deg; ++j) {
    window |= (uint64_t)(__popcll(r & state) & 1) << j;
    const bool carry = (r & top) != 0;
    r = (deg == 64) ? (r << 1) : ((r << 1) & ((1ull << deg) - 1ull));
    if (carry) r ^= poly_low;
  }
  // emit: window holds x[n .. n+deg) for the current n.  A thread owns 512 consecutive bits, so storing
  // its bytes directly would scatter every warp store over 32 lines; the packed bits go to shared memory
  // (16 per half-word) and the CTA then expands them with fully coalesced 16-byte stores.
  const size_t end = (n0 + kMseqBitsPerThread < len) ? n0 + kMseqBitsPerThread : len;
  // When every tap reaches back at least 16 elements (LTE x1: 28 and 31), the next 16 elements depend only
  // on elements already in the window: x[n+i] = XOR_t x[n+i-back_t] for i < 16 is the XOR of the window
  // shifted by (deg - back_t) - sixteen sequence elements per step instead of one.
  int top_tap = 63;
  while (top_tap > 0 && !((poly_low >> top_tap) & 1ull)) --top_tap;       // deg - (smallest back offset)
  const bool wide = deg >= 16 && deg - top_tap >= 16;
  for (int g16 = 0; g16 < kMseqBitsPerThread / 16; ++g16) {
    uint32_t b16;
    if (wide) {
      b16 = (uint32_t)window & 0xffffu;
      uint64_t nw = 0;
      for (uint64_t m = poly_low; m; m &= m - 1) nw ^= window >> (__ffsll((long long)m) - 1);
      window = (window >> 16) | ((nw & 0xffffull) << (deg - 16));
    } else {
      b16 = 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        b16 |= (uint32_t)(window & 1ull) << i;
        const uint64_t nb = (uint64_t)(__popcll(window & poly_low) & 1);
        window = (window >> 1) | (nb << (deg - 1));
      }
    }
    s_bits[g16 * kMseqPitch + threadIdx.x] = (uint16_t)b16;   // [group][thread], pitch 130: conflict-free both ways
    if (n0 + 16 * (size_t)(g16 + 1) >= end) break;       // nothing of this thread's run lies beyond len
  }
}

__global__ void __launch_bounds__(kMseqThreads) mseq_kernel(uint64_t state, uint64_t poly_low, int deg, size_t len, uint8_t* __restrict__ out,
                                                            const __grid_constant__ MseqTables tab) {
  __shared__ uint16_t s_bits[(kMseqBitsPerThread / 16) * kMseqPitch];
  mseq_fill_bits(state, poly_low, deg, len, tab, s_bits);
  __syncthreads();
  const size_t block_start = (size_t)blockIdx.x * kMseqThreads * kMseqBitsPerThread;
  const bool vec32 = ((uintptr_t)out % 32) == 0;
  for (int q2 = threadIdx.x; q2 < kMseqThreads * kMseqBitsPerThread / 32; q2 += kMseqThreads) {   // 32 bits -> 32 bytes per step
    const size_t n = block_start + 32 * (size_t)q2;
    if (n >= len) break;
    uint32_t b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int q = 2 * q2 + h;
      b[h] = s_bits[(q % (kMseqBitsPerThread / 16)) * kMseqPitch + q / (kMseqBitsPerThread / 16)];
    }
    if (n + 32 <= len && vec32) {
      const uint4 lo = make_uint4(spread4(b[0]), spread4(b[0] >> 4), spread4(b[0] >> 8), spread4(b[0] >> 12));
      const uint4 hi = make_uint4(spread4(b[1]), spread4(b[1] >> 4), spread4(b[1] >> 8), spread4(b[1] >> 12));
      st_stream_256(out + n, make_float4(__uint_as_float(lo.x), __uint_as_float(lo.y), __uint_as_float(lo.z), __uint_as_float(lo.w)),
                    make_float4(__uint_as_float(hi.x), __uint_as_float(hi.y), __uint_as_float(hi.z), __uint_as_float(hi.w)));
    } else {
      for (int i = 0; i < 32 && n + i < len; ++i) out[n + i] = (uint8_t)((b[i >> 4] >> (i & 15)) & 1u);
    }
  }
}
static uint64_t host_gf2_mulmod(uint64_t a, uint64_t b, uint64_t poly_low, int deg) {
  uint64_t r = 0;
  const uint64_t top = 1ull << (deg - 1);
  for (int i = 0; i < deg; ++i) {
    if ((b >> i) & 1ull) r ^= a;
    const bool carry = (a & top) != 0;
    a = (deg == 64) ? (a << 1) : ((a << 1) & ((1ull << deg) - 1ull));
    if (carry) a ^= poly_low;
  }
  return r;
}
void launch_mseq(uint64_t state, uint64_t poly_low, int deg, size_t len, uint8_t* out, cudaStream_t st) {
  if (len == 0) return;
  MseqTables tab;
  uint64_t sq = (deg == 1) ? poly_low : 2ull;  // z mod p
  for (int i = 0; i < 64; ++i) {
    tab.pow2[i] = sq;
    sq = host_gf2_mulmod(sq, sq, poly_low, deg);
  }
  // z^(512 t): 512 = 2^9
  uint64_t step = tab.pow2[9], r = 1ull;
  for (int t = 0; t < kMseqThreads; ++t) {
    tab.thread[t] = r;
    r = host_gf2_mulmod(r, step, poly_low, deg);
  }
  mseq_kernel<<<cdiv(len, (size_t)kMseqThreads * kMseqBitsPerThread), kMseqThreads, 0, st>>>(state, poly_low, deg, len, out, tab);
}

// =================================================================================================
// K13 statistics partial sums (the operands of the only cross-GPU reduction)
// =================================================================================================
__global__ void __launch_bounds__(256) bit_errors_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n,
                                                         ae_stats* stats, int vec_ok) {
  unsigned long long errs = 0;
  const size_t tid = (size_t)blockIdx.x * 256 + threadIdx.x, nth = (size_t)gridDim.x * 256;
  size_t done = 0;
  if (vec_ok) {
    const size_t nv = n / 16;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    const uint4* b4 = reinterpret_cast<const uint4*>(b);
    for (size_t i = tid; i < nv; i += nth) {
      const uint4 x = __ldcs(a4 + i), y = __ldcs(b4 + i);
      errs += __popc(__vcmpne4(x.x, 0) ^ __vcmpne4(y.x, 0)) + __popc(__vcmpne4(x.y, 0) ^ __vcmpne4(y.y, 0)) +
              __popc(__vcmpne4(x.z, 0) ^ __vcmpne4(y.z, 0)) + __popc(__vcmpne4(x.w, 0) ^ __vcmpne4(y.w, 0));
    }
    errs >>= 3;  // 8 mask bits per differing byte
    done = nv * 16;
  }
  for (size_t i = done + tid; i < n; i += nth) errs += ((a[i] != 0) != (b[i] != 0));
  errs = warp_sum_u64(errs);
  if ((threadIdx.x & 31) == 0 && errs) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->bit_errors), errs);
  if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(&stats->n_bits), (unsigned long long)n);
}
void launch_bit_errors(const uint8_t* a, const uint8_t* b, size_t n, ae_stats* stats, int sm_count, cudaStream_t st) {
  if (n == 0) return;
  const int vec_ok = ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0;
  unsigned g = cdiv(cdiv(n, 16), 256);
  const unsigned cap = (unsigned)sm_count * 8;
  if (g > cap) g = cap;
  bit_errors_kernel<<<g, 256, 0, st>>>(a, b, n, stats, vec_ok);
}
__global__ void __launch_bounds__(256) evm_acc_kernel(const float2* __restrict__ act, const float2* __restrict__ ref, size_t n, ae_stats* stats) {
  double e = 0.0, r = 0.0;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const float2 a = __ldcs(act + i), b = __ldcs(ref + i);
    const float dr = a.x - b.x, di = a.y - b.y;
    e += (double)(dr * dr + di * di);
    r += (double)(b.x * b.x + b.y * b.y);
  }
  e = warp_sum_f64(e); r = warp_sum_f64(r);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&stats->err_pow, e); atomicAdd(&stats->ref_pow, r); }
}
void launch_evm_acc(const float2* act, const float2* ref, size_t n, ae_stats* stats, int sm_count, cudaStream_t st) {
  if (n == 0) return;
  unsigned g = cdiv(n, 256 * 4);
  const unsigned cap = (unsigned)sm_count * 8;
  if (g > cap) g = cap;
  evm_acc_kernel<<<g, 256, 0, st>>>(act, ref, n, stats);
}

// =================================================================================================
// K15 VecStats (README.md:90-92 TODO "VecStats (f32,cf32): Min(index), Max(index), Mean(index), Power";
// SURVEY 8(f) rank 2).  One read of the vector (8 B/sample cf32, 4 B/sample f32).  cf32 elements are
// ranked by norm_sqr = re*re + im*im (unfused f32, like num-complex); ties keep the FIRST index, NaN
// never wins.  Sums are f64.  Two launches: per-CTA partials, then one CTA folds them in a fixed
// order, so the result does not depend on scheduling.
// =================================================================================================
struct StatPart {
  double sre, sim, spow;
  float mn, mx;
  unsigned long long imn, imx;
};
constexpr unsigned long long STAT_NONE = ~0ull;

__device__ __forceinline__ void stat_min(float& bv, unsigned long long& bi, float v, unsigned long long i) {
  if (i == STAT_NONE) return;
  if (bi == STAT_NONE || v < bv || (v == bv && i < bi)) { bv = v; bi = i; }
}
__device__ __forceinline__ void stat_max(float& bv, unsigned long long& bi, float v, unsigned long long i) {
  if (i == STAT_NONE) return;
  if (bi == STAT_NONE || v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
}
__device__ __forceinline__ void stat_merge(StatPart& a, const StatPart& b) {
  a.sre += b.sre; a.sim += b.sim; a.spow += b.spow;
  stat_min(a.mn, a.imn, b.mn, b.imn);
  stat_max(a.mx, a.imx, b.mx, b.imx);
}
__device__ __forceinline__ StatPart stat_shfl(const StatPart& p, int o) {
  StatPart q;
  q.sre = __shfl_xor_sync(0xffffffffu, p.sre, o);
  q.sim = __shfl_xor_sync(0xffffffffu, p.sim, o);
  q.spow = __shfl_xor_sync(0xffffffffu, p.spow, o);
  q.mn = __shfl_xor_sync(0xffffffffu, p.mn, o);
  q.mx = __shfl_xor_sync(0xffffffffu, p.mx, o);
  q.imn = __shfl_xor_sync(0xffffffffu, p.imn, o);
  q.imx = __shfl_xor_sync(0xffffffffu, p.imx, o);
  return q;
}
// fold the CTA's 256 partials; thread 0 returns the result
__device__ __forceinline__ StatPart stat_block_fold(StatPart p) {
  __shared__ StatPart sh[8];
#pragma unroll
  for (int o = 16; o; o >>= 1) { const StatPart q = stat_shfl(p, o); stat_merge(p, q); }
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    p = sh[0];
    for (int w = 1; w < 8; ++w) stat_merge(p, sh[w]);
  }
  return p;
}

// Per thread: elements i0 + it*stride, it = 0, 1, ...  The streaming loop keeps (min, max) with 32-bit
// iteration counters and plain `<` / `>` tests against +inf / -inf (NaN never passes either test); an
// element equal to the initial bound cannot win that way, so a thread that ends without a minimum or a
// maximum (its elements were all NaN / +-inf) re-walks its elements with the exact first-comparable rule.
template <bool CPLX>
__device__ __forceinline__ float stat_value(const void* base, size_t i, double& sre, double& sim, double& spow) {
  if (CPLX) {
    const float2 x = __ldcs((const float2*)base + i);
    sre += (double)x.x; sim += (double)x.y; spow += (double)x.x * (double)x.x + (double)x.y * (double)x.y;
    return __fadd_rn(__fmul_rn(x.x, x.x), __fmul_rn(x.y, x.y));
  } else {
    const float v = __ldcs((const float*)base + i);
    sre += (double)v; spow += (double)v * (double)v;
    return v;
  }
}

template <bool CPLX>
__global__ void __launch_bounds__(256) vecstats_kernel(const void* __restrict__ v, size_t n, StatPart* __restrict__ parts) {
  constexpr unsigned NONE32 = 0xffffffffu;
  const size_t stride = (size_t)gridDim.x * 256;
  const size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
  double sre = 0.0, sim = 0.0, spow = 0.0;
  float mn = __int_as_float(0x7f800000), mx = __int_as_float(0xff800000);
  unsigned itmn = NONE32, itmx = NONE32, it = 0;
  size_t i = i0;
  for (; i + 3 * stride < n; i += 4 * stride, it += 4) {   // four independent loads in flight
    float a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = stat_value<CPLX>(v, i + u * stride, sre, sim, spow);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (a[u] < mn) { mn = a[u]; itmn = it + u; }
      if (a[u] > mx) { mx = a[u]; itmx = it + u; }
    }
  }
  for (; i < n; i += stride, ++it) {
    const float a = stat_value<CPLX>(v, i, sre, sim, spow);
    if (a < mn) { mn = a; itmn = it; }
    if (a > mx) { mx = a; itmx = it; }
  }
  StatPart p{sre, sim, spow, mn, mx, itmn == NONE32 ? STAT_NONE : i0 + (size_t)itmn * stride,
             itmx == NONE32 ? STAT_NONE : i0 + (size_t)itmx * stride};
  if ((itmn == NONE32 || itmx == NONE32) && i0 < n) {       // rare: only NaN / infinities seen by this thread
    p.imn = p.imx = STAT_NONE;
    for (size_t j = i0; j < n; j += stride) {
      double d0 = 0, d1 = 0, d2 = 0;
      const float a = stat_value<CPLX>(v, j, d0, d1, d2);
      if (a == a) {
        if (p.imn == STAT_NONE || a < p.mn) { p.mn = a; p.imn = j; }
        if (p.imx == STAT_NONE || a > p.mx) { p.mx = a; p.imx = j; }
      }
    }
  }
  p = stat_block_fold(p);
  if (threadIdx.x == 0) parts[blockIdx.x] = p;
}
__global__ void __launch_bounds__(256) vecstats_fold_kernel(const StatPart* __restrict__ parts, unsigned nparts, size_t n, ae_vecstats* out) {
  StatPart p{0.0, 0.0, 0.0, 0.f, 0.f, STAT_NONE, STAT_NONE};
  for (unsigned k = threadIdx.x; k < nparts; k += 256) stat_merge(p, parts[k]);
  p = stat_block_fold(p);
  if (threadIdx.x == 0) {
    out->n = n;
    out->min_idx = p.imn == STAT_NONE ? n : p.imn;
    out->max_idx = p.imx == STAT_NONE ? n : p.imx;
    out->min_val = p.mn; out->max_val = p.mx;
    out->sum_re = p.sre; out->sum_im = p.sim; out->sum_pow = p.spow;
  }
}

size_t vecstats_scratch_bytes(int sm_count) { return (size_t)sm_count * 8 * sizeof(StatPart) + sizeof(ae_vecstats); }
int launch_vecstats(const void* v, size_t n, bool cplx, void* scratch, int sm_count, cudaStream_t st) {
  unsigned g = cdiv(n, 256 * 8);
  const unsigned cap = (unsigned)sm_count * 8;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  StatPart* parts = (StatPart*)scratch;
  ae_vecstats* out = (ae_vecstats*)((char*)scratch + (size_t)sm_count * 8 * sizeof(StatPart));
  if (cplx) vecstats_kernel<true><<<g, 256, 0, st>>>(v, n, parts);
  else vecstats_kernel<false><<<g, 256, 0, st>>>(v, n, parts);
  vecstats_fold_kernel<<<1, 256, 0, st>>>(parts, g, n, out);
  return 2;
}

}  // namespace ae
