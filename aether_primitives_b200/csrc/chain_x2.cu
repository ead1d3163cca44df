// chain_x2.cu — kernels, host tables and launcher of K14b (chain_x2.cuh): the headline chain for
// N = 1024 and <= 64 taps with packed FP32 butterflies, one warp per frame and the wrap-around
// correction on the tensor cores.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "async_copy.cuh"
#include "chain_x2.cuh"
#include "chain_x2_host.h"
#include "internal.h"

namespace ae {

// WARPS >= 14 builds are LEAN (first-stage twiddles from shared memory, twiddle products recomputed): they have to fit
// 65536 / (32 WARPS) registers
template <int N, bool INV, bool STAGED, int WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1)
chain_x2_kernel(const __grid_constant__ ChainX2Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const X2Launch L{(int)threadIdx.x, (int)blockIdx.x, (int)gridDim.x, (int)blockDim.x};
  chain_x2_body<N, INV, STAGED, (WARPS >= 14)>(p, L, smem_raw);
}

bool chain_x2_supported(size_t nfft, size_t ntaps) { return nfft == 1024 && ntaps >= 1 && ntaps <= (size_t)X2Cfg<1024>::MAX_TAPS; }

// host tables: chain_x2_host.h
void chain_x2_tables(size_t nfft, const float2* taps, size_t ntaps, std::vector<float2>& tw, std::vector<float2>& hi,
                     std::vector<float2>& lo) {
  chain_x2_twiddles(nfft, tw);
  chain_x2_split_taps(taps, ntaps, hi, lo);
}

// launcher ---------------------------------------------------------------------------------------------
template <bool INV, bool STAGED, int WARPS>
static void launch_x2(const ChainX2Params& p, cudaStream_t st) {
  using XC = X2Cfg<1024>;
  auto kern = chain_x2_kernel<1024, INV, STAGED, WARPS>;
  const size_t smem = XC::smem_bytes(WARPS, STAGED);
  static thread_local int cached_dev = -1, sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  if (cached_dev != dev) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  const size_t want = (p.frames + WARPS - 1) / WARPS;      // one frame per warp at least
  const unsigned grid = (unsigned)(want < (size_t)sms ? want : (size_t)sms);
  kern<<<grid, 32 * WARPS, smem, st>>>(p);
}

template <bool INV>
static void launch_x2_dir(const ChainX2Params& p, cudaStream_t st) {
  static const char* no_tma = getenv("AE_CHAIN_NO_TMA");
  static const char* warps_env = getenv("AE_CHAIN_WARPS");
  const bool staged = ((uintptr_t)p.x % 16) == 0 && !no_tma;
  // 8 warps (2 per sub-partition) measured best: the kernel is bound by instruction dispatch, not by latency
  const int warps = warps_env ? atoi(warps_env) : (staged ? 8 : 16);   // multiples of 4 only: 10 or 14 warps leave the sub-partitions unbalanced (measured 12 % slower)
  if (staged) {
    if (warps <= 4) launch_x2<INV, true, 4>(p, st);
    else if (warps <= 8) launch_x2<INV, true, 8>(p, st);
    else launch_x2<INV, true, 10>(p, st);
  } else {
    if (warps <= 8) launch_x2<INV, false, 8>(p, st);
    else if (warps <= 12) launch_x2<INV, false, 12>(p, st);
    else if (warps <= 14) launch_x2<INV, false, 14>(p, st);
    else if (warps <= 16) launch_x2<INV, false, 16>(p, st);
    else launch_x2<INV, false, 20>(p, st);
  }
}

void launch_chain_x2(const float2* x, uint8_t* bits, size_t frames, const float2* window, const float2* tw, const float2* taps_hi,
                     const float2* taps_lo, size_t ntaps, bool inverse, float scale, int compat, cudaStream_t st) {
  if (frames == 0) return;
  ChainX2Params p;
  p.x = x; p.bits = bits; p.frames = frames; p.window = window; p.tw = tw; p.taps_hi = taps_hi; p.taps_lo = taps_lo;
  p.ntaps = (int)ntaps; p.scale = scale; p.compat = compat;
  static const char* dbg = getenv("AE_CHAIN_DEBUG");
  p.debug = dbg ? atoi(dbg) : 0;
  static const char* stg = getenv("AE_CHAIN_STAGGER");
  p.stagger = stg ? atoi(stg) : 2400;
  if (inverse) launch_x2_dir<true>(p, st);
  else launch_x2_dir<false>(p, st);
}

}  // namespace ae
