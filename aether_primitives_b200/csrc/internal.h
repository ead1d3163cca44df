// internal.h — launcher interface between the host API (api.cu) and the kernel TUs.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <vector>
#include "../../include/aether_b200.h"

namespace ae {

// Resident CTAs (SM count x CTAs per SM) of `kern` at (threads, dynamic shared memory) on the current device,
// including the opt-in for more than 48 KB of dynamic shared memory.  Queried once per (kernel, device, size): the
// reference API is one frame per call, where three driver calls per launch would be most of the cost.
size_t resident_ctas(const void* kern, int threads, size_t smem);
// Launchers are void; an unsupported length / shape is recorded here and turned into AE_EARG by the next CKL() in
// api.cu instead of silently launching nothing.
void note_unsupported_launch(const char* what);
const char* take_unsupported_launch();

// ---- K1 fused VecOps tape -------------------------------------------------------------------
enum TapeOpcode : int { OP_SCALE = 0, OP_MUL, OP_DIV, OP_ADD, OP_SUB, OP_CONJ, OP_MIRROR, OP_CLONE, OP_ZERO };
constexpr int kMaxTape = 16;
struct TapeEntry {
  const float2* operand;  // other[] for MUL/DIV/ADD/SUB/CLONE
  float s;                // SCALE factor
  int op;
};
struct TapeParams {
  TapeEntry e[kMaxTape];
  int n_ops;
  int load_self;  // 0 when a ZERO/CLONE precedes every use of the old contents
};
void launch_vecops(float2* v, size_t n, const TapeParams& p, bool has_mirror, int sm_count, cudaStream_t st);

// ---- K5/K6 sampling --------------------------------------------------------------------------
void launch_downsample_cf32(const float2* src, float2* dst, size_t n_dst, size_t dec, cudaStream_t st);
void launch_downsample_u8(const uint8_t* src, uint8_t* dst, size_t n_dst, size_t dec, cudaStream_t st);
void launch_interpolate(const float2* src, size_t n_src, float2* dst, size_t n_between, int compat, cudaStream_t st);

// ---- K7/K8 modulation ------------------------------------------------------------------------
struct ModTable { float2 t[4]; int len; int generic_qpsk; };
void launch_modulate(const ModTable& tab, const uint8_t* bits, size_t nbits, float2* out, size_t n_out,
                     int* errflag, cudaStream_t st);
void launch_demod(const ModTable& tab, const float2* sym, size_t n, uint8_t* bits, int compat, cudaStream_t st);

// ---- K9/K10 noise ----------------------------------------------------------------------------
void launch_awgn_fill(float2* dst, size_t n, float scale, uint64_t seed, uint64_t stream, uint64_t offset, cudaStream_t st);
void launch_awgn_apply(float2* sig, size_t n, float scale, int twice, uint64_t seed, uint64_t stream, uint64_t offset,
                       cudaStream_t st);
void launch_modem_fused(const ModTable& tab, const uint8_t* bits_in, size_t nbits, uint8_t* bits_out, float scale,
                        int twice, uint64_t seed, uint64_t stream, uint64_t offset, int compat, ae_stats* stats,
                        int* errflag, int sm_count, cudaStream_t st);

// ---- K11 sequence ----------------------------------------------------------------------------
void launch_expand(uint64_t seed, size_t len, uint8_t* out, cudaStream_t st);
// state = first `deg` bits of the recurrence window (bit i = x[base+i]); poly_low = coefficients of
// x^0..x^(deg-1) of the characteristic polynomial; writes out[0..len) = x[base + 0 ...]
void launch_mseq(uint64_t state, uint64_t poly_low, int deg, size_t len, uint8_t* out, cudaStream_t st);

// ---- K13 statistics --------------------------------------------------------------------------
void launch_bit_errors(const uint8_t* a, const uint8_t* b, size_t n, ae_stats* stats, int sm_count, cudaStream_t st);
void launch_evm_acc(const float2* act, const float2* ref, size_t n, ae_stats* stats, int sm_count, cudaStream_t st);
// VecStats: scratch = per-CTA partials followed by the ae_vecstats result; returns the launch count
size_t vecstats_scratch_bytes(int sm_count);
int launch_vecstats(const void* v, size_t n, bool cplx, void* scratch, int sm_count, cudaStream_t st);

// ---- K2 FFT ----------------------------------------------------------------------------------
// power-of-two register/shared-memory kernel, 16 <= n <= 16384.  `tw` of every power-of-two launcher
// below (FFT, overlap-save FIR, chains) is the PER-THREAD table built by fft_thread_twiddles().
bool fft_pow2_supported(size_t n);
void fft_thread_twiddles(size_t n, std::vector<float2>& out);
void launch_fft_pow2(const float2* in, float2* out, size_t n, size_t frames, const float2* tw, bool inverse,
                     bool do_scale, float scale, cudaStream_t st);
// 2^15..2^24: four-step, two column-FFT passes; tw1/tw2 = per-thread tables of the two factor lengths,
// wlo[e] = exp(-2 pi i e/n) (e < 4096), whi[e] = exp(-2 pi i 4096 e/n)
bool fft_big_supported(size_t n);
void fft_big_split(size_t n, size_t* n1, size_t* n2);
void launch_fft_big(const float2* in, float2* out, float2* scratch, size_t n, size_t frames, const float2* tw1, const float2* tw2,
                    const float2* wlo, const float2* whi, bool inverse, bool do_scale, float scale, cudaStream_t st);
// any length: one global-memory Stockham pass per prime-power factor; needs 2 scratch buffers
// Bluestein helpers (any length through the power-of-two kernels); log2m = log2 of the padded length
void launch_bluestein_pre(const float2* in, float2* a, const float2* w, size_t n, unsigned log2m, size_t frames, bool conj_in, int sm_count,
                          cudaStream_t st);
void launch_bluestein_mul(float2* a, const float2* bspec, unsigned log2m, size_t frames, int sm_count, cudaStream_t st);
void launch_bluestein_post(const float2* c, float2* out, const float2* w, size_t n, unsigned log2m, size_t frames, bool conj_out,
                           bool do_scale, float scale, cudaStream_t st);
void launch_fft_generic(const float2* in, float2* out, float2* scratch, size_t n, size_t frames, const float2* tw,
                        const uint32_t* radices, int n_radices, bool inverse, bool do_scale, float scale,
                        cudaStream_t st);
// ---- K3/K4 FIR -------------------------------------------------------------------------------
void launch_fir_direct(const float2* x, float2* y, size_t n, const float2* taps_padded, int ntaps_padded,
                       const float2* history /*ntaps_padded-1 samples or null*/, size_t frame_len, int sm_count,
                       cudaStream_t st, const float2* taps_host_padded = nullptr /* host copy: short filters keep their taps in the constant bank */);
bool fir_os_supported(size_t nfft);
// H = FFT_nfft(taps)/nfft (exp(-) convention); L = nfft - ntaps + 1 outputs per segment
void launch_fir_overlap_save(const float2* x, float2* y, size_t n, const float2* H, const float2* tw, size_t nfft,
                             size_t ntaps, const float2* history, size_t frame_len, cudaStream_t st);
// K4b: 1024-point blocks on the packed one-exchange transform of chain_x2.cuh; x2tw = chain_x2_twiddles rows
void launch_fir_os_x2(const float2* x, float2* y, size_t n, const float2* H, const float2* x2tw, size_t ntaps, const float2* history,
                      size_t frame_len, cudaStream_t st);
// ---- K14 / K12 chains ------------------------------------------------------------------------
bool chain_fused_supported(size_t nfft, size_t ntaps);
// window[n] = s * sum_k h[k] exp(-sgn 2 pi i nk/N); taps in device memory
void launch_chain_fused(const float2* x, uint8_t* bits, size_t nfft, size_t frames, const float2* window,
                        const float2* taps, size_t ntaps, const float2* tw, bool inverse, float scale, int compat,
                        cudaStream_t st);
// K14b (chain_x2.cuh): N = 1024, <= 64 taps; packed FP32 butterflies, warp per frame, tensor-core fix-up.
// tw = chain_x2_twiddles() rows, taps_hi/lo = chain_x2_split_taps(); bits must be 4-byte aligned.
bool chain_x2_supported(size_t nfft, size_t ntaps);
void chain_x2_tables(size_t nfft, const float2* taps, size_t ntaps, std::vector<float2>& tw, std::vector<float2>& hi,
                     std::vector<float2>& lo);
void launch_chain_x2(const float2* x, uint8_t* bits, size_t frames, const float2* window, const float2* tw, const float2* taps_hi,
                     const float2* taps_lo, size_t ntaps, bool inverse, float scale, int compat, cudaStream_t st);
bool ofdm_supported(size_t nfft);
void launch_ofdm_chain(size_t nfft, size_t frames, uint64_t first_frame, float noise_scale, int twice, uint64_t seed,
                       const float2* tw, int compat, uint8_t* tx_bits, uint8_t* rx_bits, ae_stats* stats,
                       cudaStream_t st);

// ---- spectral: spectrogram core and frequency-domain correlator --------------------
bool spectral_supported(size_t n);
void launch_spectrogram(const float2* in, size_t n_samples, float* levels, size_t n, size_t frames, const float2* tw, bool inverse,
                        float scale, int use_db, cudaStream_t st);
void launch_levels(const float2* spec, float* levels, size_t total, size_t n, int use_db, cudaStream_t st);
// 1024-point frames on the packed one-exchange transform (warp per frame); x2tw = chain_x2_twiddles rows
void launch_correlate_x2(float2* data, const float2* sig, size_t frames, const float2* x2tw, bool fwd_inverse, float scale, int do_scale,
                         cudaStream_t st);
void launch_correlate(float2* data, const float2* sig, size_t n, size_t frames, const float2* tw, bool fwd_inverse, float scale,
                      int do_scale, cudaStream_t st);

}  // namespace ae
