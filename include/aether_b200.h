/* aether_b200.h — C ABI of the B200-native cf32 signal-processing path.
 *
 * This is the drop-in boundary underneath the reference crate's Rust traits
 * (razorheadfx/aether_primitives).  The reference has no FFI of its own; each entry
 * point below names the reference interface it replaces (file:line relative to the
 * reference repository).  INTEGRATION.md shows the Rust `extern "C"` binding a
 * maintainer would add (VecOps for DeviceVec, Fft for CudaFft, Modulation for
 * DeviceTable).
 *
 * Conventions
 *  - every function returns an ae_status (0 = AE_OK); nothing throws or aborts across the
 *    ABI.  The reference panics instead (assert_eq!/unwrap); the host-side mirror maps a
 *    non-zero status to the reference's panic message.  ae_last_error_string() gives it.
 *  - cf32 = two f32 back to back (src/lib.rs:8-12); bits are one u8 per bit
 *    (src/modulation.rs:102-103).
 *  - threading: the per-device context (pending op tapes, table caches, the current stream) is not
 *    locked.  All calls that touch one device, whatever the handle, must come from one thread
 *    at a time; different devices may be driven from different threads.  Error strings and the
 *    current device (ae_init) are per thread.  All work of a device is issued on one CUDA
 *    stream (ae_set_stream), so calls are ordered as issued; stage pipelines (ae_pipeline_*) get
 *    their concurrency from CUDA streams, not from host threads.
 *  - VecOps calls on an ae_vec are RECORDED on the handle's op tape and executed as ONE
 *    fused elementwise kernel when a consumer needs the data (ae_vec_flush, download, FFT,
 *    FIR, use as operand, free).  Results are visible after ae_sync()/download.
 *  - There is no CPU fallback: without a CUDA device every compute call returns AE_ECUDA.
 */
#ifndef AETHER_B200_H
#define AETHER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ae_cf32 { float re, im; } ae_cf32;   /* src/lib.rs:12 */

typedef int ae_status;
enum {
  AE_OK = 0,
  AE_ELEN = 1,   /* "Vectors must have same length" / "Input and FFT must be the same length" */
  AE_EARG = 2,   /* invalid argument (null handle, zero length where Rust would panic, ...) */
  AE_ECUDA = 3,  /* CUDA runtime error (sticky; see ae_last_error_string) */
  AE_ENCCL = 4,
  AE_EOOM = 5,
  AE_EIDX = 6    /* index out of bounds where the reference panics on slice indexing */
};

/* behaviour switch for the reference quirks SURVEY.md F3/F4/F5 */
enum { AE_COMPAT_REFERENCE = 0, AE_COMPAT_CORRECTED = 1 };

/* fft::Scale (src/fft.rs:6-18) */
enum { AE_SCALE_NONE = 0, AE_SCALE_SN = 1, AE_SCALE_N = 2, AE_SCALE_X = 3 };

typedef struct ae_vec ae_vec;     /* device Vec<cf32> / &mut [cf32] */
typedef struct ae_bits ae_bits;   /* device Vec<u8>, one byte per bit */
typedef struct ae_fft ae_fft;     /* fft::Cfft */
typedef struct ae_fir ae_fir;     /* fir::Fir */
typedef struct ae_mod ae_mod;     /* impl Modulation for [cf32; 2|4] */
typedef struct ae_awgn ae_awgn;   /* noise::Awgn */
typedef struct ae_chain ae_chain; /* fused FFT -> FIR -> QPSK demod plan */

/* {bit_errors, n_bits, sum|e|^2, sum|r|^2}: the only quantities ever reduced across GPUs */
typedef struct ae_stats { uint64_t bit_errors, n_bits; double err_pow, ref_pow; } ae_stats;

/* ---- runtime ------------------------------------------------------------------------ */
const char* ae_version(void);
ae_status ae_device_count(int* n);
ae_status ae_init(int device);                 /* cudaSetDevice + per-device context */
ae_status ae_set_stream(void* cuda_stream);    /* NULL: the library's own stream */
void*     ae_get_stream(void);
ae_status ae_sync(void);                       /* stream sync + deferred device-side errors */
const char* ae_last_error_string(void);
ae_status ae_sm_count(int* n);
uint64_t  ae_launch_count(void);               /* kernels launched by this library so far */
ae_status ae_host_alloc(size_t bytes, void** p);   /* pinned host memory for the *_host entry points */
ae_status ae_host_free(void* p);

/* ---- device Vec<cf32> (src/vecops.rs:331-332 impl targets [cf32], Vec<cf32>) ---------- */
ae_status ae_vec_alloc(size_t len, size_t capacity, ae_vec** out);   /* vec![0; len] / with_capacity */
ae_status ae_vec_wrap(void* device_ptr, size_t len, ae_vec** out);   /* borrow foreign device memory */
ae_status ae_vec_view(ae_vec* parent, size_t offset, size_t len, ae_vec** out); /* &mut v[a..b] */
ae_status ae_vec_free(ae_vec* v);
size_t    ae_vec_len(const ae_vec* v);
size_t    ae_vec_capacity(const ae_vec* v);
ae_status ae_vec_set_len(ae_vec* v, size_t len);                     /* <= capacity */
ae_status ae_vec_reserve(ae_vec* v, size_t capacity);
ae_status ae_vec_device_ptr(ae_vec* v, void** ptr);                  /* flushes the tape */
ae_status ae_vec_upload(ae_vec* v, const ae_cf32* host, size_t n);   /* n == len */
ae_status ae_vec_download(ae_vec* v, ae_cf32* host, size_t n);       /* flushes, synchronises */

/* VecOps (src/vecops.rs:39-89).  Recorded on the tape; return AE_ELEN when the reference's
 * assert_eq!(self.len(), other.len()) would panic (:100-104 etc). */
ae_status ae_vec_scale(ae_vec* v, float s);                  /* :94-97  */
ae_status ae_vec_mul(ae_vec* v, ae_vec* other);              /* :99-112 */
ae_status ae_vec_div(ae_vec* v, ae_vec* other);              /* :114-125 */
ae_status ae_vec_conj(ae_vec* v);                            /* :127-130 */
ae_status ae_vec_add(ae_vec* v, ae_vec* other);              /* :132-142 */
ae_status ae_vec_sub(ae_vec* v, ae_vec* other);              /* :144-155 */
ae_status ae_vec_mirror(ae_vec* v);                          /* :157-161 */
ae_status ae_vec_clone(ae_vec* v, ae_vec* other);            /* :163-172 */
ae_status ae_vec_zero(ae_vec* v);                            /* :174-177 */
/* vec_mutate (:179-182): arbitrary closure, element order -> host round trip (slow path) */
ae_status ae_vec_mutate(ae_vec* v, void (*f)(ae_cf32* elem, void* user), void* user);
ae_status ae_vec_flush(ae_vec* v);                           /* run the pending tape now */
size_t    ae_vec_pending_ops(const ae_vec* v);
/* Scale::scale (src/fft.rs:22-37) */
ae_status ae_scale_factor(int scale_kind, size_t n, float x, float* s);
ae_status ae_vec_scale_kind(ae_vec* v, int scale_kind, float x);
/* vec_fft / vec_ifft (src/vecops.rs:301-313): plan on the fly, in place, one frame = whole vec */
ae_status ae_vec_fft(ae_vec* v, int scale_kind, float x, int compat);
ae_status ae_vec_ifft(ae_vec* v, int scale_kind, float x, int compat);

/* ---- device Vec<u8> ------------------------------------------------------------------- */
ae_status ae_bits_alloc(size_t len, size_t capacity, ae_bits** out);
ae_status ae_bits_wrap(void* device_ptr, size_t len, ae_bits** out);
ae_status ae_bits_free(ae_bits* b);
size_t    ae_bits_len(const ae_bits* b);
size_t    ae_bits_capacity(const ae_bits* b);
ae_status ae_bits_set_len(ae_bits* b, size_t len);
ae_status ae_bits_device_ptr(ae_bits* b, void** ptr);
ae_status ae_bits_upload(ae_bits* b, const uint8_t* host, size_t n);
ae_status ae_bits_download(ae_bits* b, uint8_t* host, size_t n);

/* ---- fft::Fft / fft::Cfft (src/fft.rs:48-77, :134-235) --------------------------------- */
enum { AE_FFT_FWD = 0, AE_FFT_BWD = 1 };
ae_status ae_fft_create(size_t len, ae_fft** out);           /* Cfft::with_len :147 (any len >= 1) */
ae_status ae_fft_destroy(ae_fft* f);
size_t    ae_fft_len(const ae_fft* f);                       /* Fft::len :232 */
/* compat=reference: "fwd" uses exp(+2 pi i nk/N) exactly like the crate */
ae_status ae_fft_set_compat(ae_fft* f, int compat);
/* fwd/bwd (:162-182): out != in, input preserved.  ifwd/ibwd (:184-204): out == NULL.
 * `howmany` frames of len() samples (batching is new surface; howmany = 1 is the reference
 * call).  AE_ELEN unless in.len == len*howmany (and out.len likewise). */
ae_status ae_fft_exec(ae_fft* f, int dir, ae_vec* in, ae_vec* out, int scale_kind, float x,
                      size_t howmany);
/* tfwd/tbwd (:206-230): result lives in the plan's scratch; *view stays valid until the next
 * call on this plan and must not be freed. */
ae_status ae_fft_exec_tmp(ae_fft* f, int dir, ae_vec* in, int scale_kind, float x, size_t howmany,
                          ae_vec** view);

/* ---- fir (src/fir.rs:3-22 is a constructor-only stub; semantics defined in DESIGN.md) -- */
enum { AE_FIR_AUTO = 0, AE_FIR_DIRECT = 1, AE_FIR_OVERLAP_SAVE = 2 };   /* AUTO = overlap-save whenever a block length exists (faster at every tap count), else direct */
ae_status ae_fir_create(const ae_cf32* taps_host, size_t ntaps, int mode, ae_fir** out); /* Fir::new :14 */
ae_status ae_fir_destroy(ae_fir* f);
size_t    ae_fir_ntaps(const ae_fir* f);
/* Outputs per overlap-save segment (FFT block length - ntaps + 1; 1 for the direct form).  A stream that is cut into
 * shards whose first sample is a multiple of this hop is filtered with exactly the segment boundaries of the whole
 * stream, so shard results are bit-identical to the unsharded run (multi-GPU FIR). */
size_t    ae_fir_block_hop(const ae_fir* fir);
ae_status ae_fir_reset(ae_fir* f);                            /* zero the carried history */
/* y[n] = sum_k h[k] x[n-k]; out.len == in.len.  frame_len = 0: one stream, history carried
 * across calls (T-1 samples); frame_len > 0: zero state at the start of every frame. */
ae_status ae_fir_exec(ae_fir* f, ae_vec* in, ae_vec* out, size_t frame_len);

/* ---- sampling (src/sampling.rs) -------------------------------------------------------- */
/* interpolate :7-24 — APPENDS (n-1)*(n_between+1)+1 samples to dst (grows it if needed). */
ae_status ae_interpolate(ae_vec* src, ae_vec* dst, size_t n_between, int compat);
/* downsample :28-42 / downsample_sb :49-62 — dst[i] = src[i*(src.len/dst.len)];
 * strict != 0 enforces the debug_assert on divisibility. */
ae_status ae_downsample(ae_vec* src, ae_vec* dst, int strict);
ae_status ae_downsample_sb(ae_vec* src, ae_vec* dst, int strict);
ae_status ae_downsample_bits(ae_bits* src, ae_bits* dst, int strict);   /* generic T: Copy */

/* ---- modulation (src/modulation.rs) ---------------------------------------------------- */
ae_status ae_mod_create(const ae_cf32* table, size_t table_len, ae_mod** out); /* 2 or 4 entries */
ae_status ae_mod_bpsk(ae_mod** out);                          /* bpsk() :61-63, table :77 */
ae_status ae_mod_qpsk(ae_mod** out);                          /* qpsk() :66-68, table :87-92 */
ae_status ae_mod_destroy(ae_mod* m);
size_t    ae_mod_bits_per_symbol(const ae_mod* m);            /* :146-148 */
/* modulate :115-121 — out is overwritten: out.len = ceil(bits.len / BPS) (collect()).
 * A ragged QPSK tail or an index >= table_len is reported as AE_EIDX by the next ae_sync. */
ae_status ae_mod_modulate(ae_mod* m, ae_bits* bits, ae_vec* out);
/* modulate_into :123-131 — writes min(symbols, out.len) symbols, never resizes out */
ae_status ae_mod_modulate_into(ae_mod* m, ae_bits* bits, ae_vec* out);
/* demod_naive :133-144 (BPSK) / :33-56 (QPSK override) — APPENDS BPS bytes per symbol */
ae_status ae_mod_demod(ae_mod* m, ae_vec* symbols, ae_bits* out, int compat);

/* ---- noise (src/noise.rs) --------------------------------------------------------------- */
ae_status ae_awgn_create(float power, uint64_t seed, ae_awgn** out);   /* noise::new :14-16 */
ae_status ae_awgn_generator(ae_awgn** out);                            /* generator() :9-11 */
ae_status ae_awgn_destroy(ae_awgn* g);
ae_status ae_awgn_set_power(ae_awgn* g, float power);                  /* :47-50 */
ae_status ae_awgn_set_stream_id(ae_awgn* g, uint64_t stream_id);       /* Philox subsequence */
ae_status ae_awgn_seek(ae_awgn* g, uint64_t sample_offset);
uint64_t  ae_awgn_tell(const ae_awgn* g);
ae_status ae_awgn_fill(ae_awgn* g, ae_vec* target);                    /* :62-66 len -> capacity */
ae_status ae_awgn_apply(ae_awgn* g, ae_vec* signal, int compat);       /* :53-59 */
/* Philox4x32-10 block function, exposed for the Random123 known-answer tests */
ae_status ae_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* ---- sequence (src/sequence.rs) --------------------------------------------------------- */
ae_status ae_mseq_expand(uint64_t seed, size_t len, ae_bits* out);     /* expand :18-21, out overwritten */
/* generate :47-53 with the generator restricted to x[n] = (sum_t x[n-back[t]]) % 2 (:42) */
ae_status ae_mseq_generate(const uint8_t* init_host, size_t n_init, const uint32_t* back_offsets,
                           size_t n_back, size_t len, ae_bits* out);

/* ---- statistics (the quantities that are all-reduced across GPUs) ----------------------- */
ae_status ae_stats_alloc(ae_stats** device_stats);            /* zeroed, device memory */
ae_status ae_stats_free(ae_stats* device_stats);
ae_status ae_stats_zero(ae_stats* device_stats);
ae_status ae_stats_read(const ae_stats* device_stats, ae_stats* host);  /* synchronises */
ae_status ae_count_bit_errors(ae_bits* a, ae_bits* b, ae_stats* device_stats);
ae_status ae_evm_accumulate(ae_vec* actual, ae_vec* reference, ae_stats* device_stats);

/* ---- the one collective of the path: sum of ae_stats over the GPUs, NCCL over NVLink.
 * The reference is single-process and has no counterpart; EVM in dB follows src/lib.rs:21 from the
 * reduced sums.  libnccl.so.2 is opened at the first call (the library itself loads without it);
 * every failure returns AE_ENCCL with ncclGetErrorString in ae_last_error_string().
 *   one process per GPU : rank 0 calls ae_comm_unique_id and hands the 128 bytes to every rank out of
 *                         band (MPI, a file, torch.distributed ...); each rank then calls ae_comm_init_rank
 *                         on its own device and ae_stats_allreduce on the context stream.
 *   one process, n GPUs : ae_comm_init_all (ncclCommInitAll on devices 0..n-1) + ae_stats_allreduce_all,
 *                         which issues every device's all-reduce inside one NCCL group. */
#define AE_COMM_ID_BYTES 128
typedef struct ae_comm ae_comm;
ae_status ae_comm_unique_id(uint8_t id_out[AE_COMM_ID_BYTES]);
ae_status ae_comm_init_rank(const uint8_t id[AE_COMM_ID_BYTES], int nranks, int rank, ae_comm** out);
ae_status ae_comm_init_all(int ndev, ae_comm** comms_out /* ndev handles */);
ae_status ae_comm_destroy(ae_comm* comm);
ae_status ae_comm_info(const ae_comm* comm, int* nranks, int* rank, int* device);
/* in place: device_stats <- sum over ranks; asynchronous on the device context's stream
 * (ae_stats_read synchronises).  u64 counters are summed exactly; the two f64 sums in NCCL's order. */
ae_status ae_stats_allreduce(ae_stats* device_stats, ae_comm* comm);
ae_status ae_stats_allreduce_all(ae_stats** device_stats, ae_comm** comms, int ndev);

/* VecStats, the README TODO "Add VecStats (f32,cf32): Min(index), Max(index), Mean(index), Power"
 * (README.md:90-92; no definition exists in the reference, so this IS the definition):
 * cf32 elements are ranked by norm_sqr = re*re + im*im (f32, unfused), f32 elements by value; ties keep
 * the first index, NaN never wins, min_idx = max_idx = n when nothing is comparable; sums are f64
 * (mean = sum/n, power = sum_pow/n).  One pass over the vector; synchronises.  Empty input: AE_ELEN. */
typedef struct ae_vecstats {
  uint64_t n, min_idx, max_idx;
  float min_val, max_val;
  double sum_re, sum_im, sum_pow;
} ae_vecstats;
ae_status ae_vec_stats(ae_vec* v, ae_vecstats* host_out);

/* ---- fused chains ----------------------------------------------------------------------- */
/* examples/modem.rs:15-32 in one pass: modulate -> Awgn::apply -> demod_naive (+ bit errors);
 * symbols never touch HBM.  bits_out is overwritten (len = bits_in.len). */
ae_status ae_modem_fused(ae_mod* m, ae_awgn* g, ae_bits* bits_in, ae_bits* bits_out,
                         ae_stats* device_stats, int compat);

/* headline chain: per frame Cfft::fwd(scale) -> FIR (zero state per frame) -> QPSK demod */
ae_status ae_chain_create(size_t fft_len, const ae_cf32* taps_host, size_t ntaps, int scale_kind,
                          float x, int compat, ae_chain** out);
ae_status ae_chain_destroy(ae_chain* c);
ae_status ae_chain_exec(ae_chain* c, ae_vec* in, ae_bits* bits_out);    /* bits_out overwritten */
/* same with HOST buffers: chunked H2D -> kernel -> D2H pipeline on internal streams */
ae_status ae_chain_exec_host(ae_chain* c, const ae_cf32* host_in, size_t n_samples,
                             uint8_t* host_bits);
/* Streaming form of the same pipeline — the analogue of src/pipeline.rs:26-137 (stages connected by
 * channels, a report per stage: processed / active time / rate / utilisation, :93-107) and of
 * src/pool.rs:43-130 (a pool of reusable buffers) for this path: the stages are H2D copy, the fused
 * kernel and D2H copy; the pool is a ring of `depth` device buffer slots.  ae_pipe_send queues one
 * block of block_frames frames and returns at once (when every slot is in flight it first waits for
 * the oldest block, like taking from an empty pool); ae_pipe_recv hands back, in order, the
 * host_bits pointer of the next finished block (waiting for it if necessary).  Host buffers should
 * be pinned (ae_host_alloc) and must stay valid until received. */
typedef struct ae_pipe ae_pipe;
typedef struct ae_pipe_stage {
  char name[16];
  uint64_t processed;                 /* blocks since the last reset */
  double active_ms, elapsed_ms;       /* union of the stage's busy intervals / report window */
  double per_second, utilisation_pct;
} ae_pipe_stage;
ae_status ae_pipe_create(ae_chain* c, size_t block_frames, int depth, ae_pipe** out);
ae_status ae_pipe_destroy(ae_pipe* p);
ae_status ae_pipe_send(ae_pipe* p, const ae_cf32* host_in, uint8_t* host_bits);
ae_status ae_pipe_recv(ae_pipe* p, uint8_t** host_bits_done);
size_t    ae_pipe_in_flight(const ae_pipe* p);      /* sent and not yet received */
ae_status ae_pipe_report(ae_pipe* p, ae_pipe_stage stages[3], int reset);

/* General form (src/pipeline.rs:26-137: pipeline::new(name, op).add_stage(name, op)...finish()): any number of named
 * stages, each with its own CUDA stream in place of the reference's thread and CUDA events in place of its channels.
 * ae_pipeline_send(item) calls every stage's `op` once, in order, on the calling thread; while `op` runs, the library's
 * context stream IS that stage's stream, so whatever the stage does through this API (kernels, ae_*_async copies) is
 * queued there and `op` returns without waiting.  Stage k of item i therefore overlaps stage k+1 of item i-1, exactly like
 * the reference's stage threads; items leave in the order they entered.  At most `depth` items are in flight (send waits
 * for the oldest one otherwise); ae_pipeline_recv returns the next finished item, waiting for its last stage.  The report
 * is the reference's per-stage line (:93-107).  Rules: a handle with internal scratch (FFT plan, FIR state, Awgn stream)
 * belongs to ONE stage; an `op` must not call anything that synchronises (downloads, ae_sync, ae_stats_read) nor
 * ae_set_stream; buffers an item carries are reused only after it has been received (take them from a pool). */
typedef struct ae_pipeline ae_pipeline;
typedef ae_status (*ae_stage_fn)(void* user, size_t slot /* 0..depth-1, stable while the item is in flight */, void* item);
ae_status ae_pipeline_create(int depth, ae_pipeline** out);
ae_status ae_pipeline_add_stage(ae_pipeline* p, const char* name, ae_stage_fn op, void* user);
ae_status ae_pipeline_send(ae_pipeline* p, void* item);
ae_status ae_pipeline_recv(ae_pipeline* p, void** item_done);
size_t    ae_pipeline_in_flight(const ae_pipeline* p);
size_t    ae_pipeline_stages(const ae_pipeline* p);
ae_status ae_pipeline_report(ae_pipeline* p, ae_pipe_stage* stages, size_t n_stages, int reset);
ae_status ae_pipeline_destroy(ae_pipeline* p);       /* waits for everything in flight */
/* stream-ordered copies for pipeline stages: no synchronisation; `host` must be pinned (ae_host_alloc) to overlap and
 * must stay valid until the item is received */
ae_status ae_vec_upload_async(ae_vec* v, const ae_cf32* host, size_t n);
ae_status ae_vec_download_async(ae_vec* v, ae_cf32* host, size_t n);
ae_status ae_bits_upload_async(ae_bits* b, const uint8_t* host, size_t n);
ae_status ae_bits_download_async(ae_bits* b, uint8_t* host, size_t n);

/* ---- CUDA graphs: launch-bound shapes (examples/modem.rs runs 1M symbols, a ~4 us kernel) ----------
 * Between ae_graph_begin and ae_graph_end every launch the library makes on the context stream is RECORDED instead of
 * run (cudaStreamBeginCapture); ae_graph_launch replays the whole recording with one driver call.  Rules of stream
 * capture apply: size every output beforehand (nothing may reallocate) and call nothing that synchronises
 * (downloads, ae_sync, ae_stats_read) inside the bracket.  Host-side state advances at record time: an Awgn stream
 * offset is baked into each recorded launch, so the k launches of one recording use k consecutive noise blocks and
 * every replay repeats them.  VecOps still pending when the bracket opens run once, before it; VecOps recorded inside it
 * and not yet run when it closes are part of the graph. */
typedef struct ae_graph ae_graph;
ae_status ae_graph_begin(void);
ae_status ae_graph_end(ae_graph** out);
ae_status ae_graph_launch(ae_graph* g);
ae_status ae_graph_destroy(ae_graph* g);
/* unfused composition of the same chain from the stand-alone kernels (for cross-checking) */
ae_status ae_chain_exec_unfused(ae_chain* c, ae_vec* in, ae_bits* bits_out, ae_vec* symbols_out);

/* OFDM-like chain (BASELINE config 5), one frame per CTA, nothing but audit bits stored:
 * M-sequence(frame seed) -> QPSK -> bwd FFT(SN) -> AWGN -> fwd FFT(SN) -> demod -> BER/EVM */
ae_status ae_ofdm_chain(size_t fft_len, size_t frames, uint64_t first_frame_id, float noise_power,
                        uint64_t noise_seed, int compat, ae_bits* tx_bits /*nullable*/,
                        ae_bits* rx_bits /*nullable*/, ae_stats* device_stats);

/* ---- SURVEY 8(f) "next" rows ----------------------------------------------------------------- */
typedef struct ae_f32 ae_f32;     /* device Vec<f32> (spectrogram levels) */
ae_status ae_f32_alloc(size_t len, ae_f32** out);
ae_status ae_f32_free(ae_f32* v);
size_t    ae_f32_len(const ae_f32* v);
ae_status ae_f32_device_ptr(ae_f32* v, void** ptr);
ae_status ae_f32_download(ae_f32* v, float* host, size_t n);
ae_status ae_f32_stats(ae_f32* v, ae_vecstats* host_out);   /* sum_im = 0 */
/* compute core of util::plot::waterfall / spectrum (src/util/plot.rs:46-68, :109-130): symbols are
 * zero padded to a multiple of fft.len(); per chunk vec_rfft(Scale::SN) -> vec_mirror -> c.norm() ->
 * DB::from(..).db() when use_db (src/util/mod.rs:26-34).  levels is resized to chunks*len. */
ae_status ae_spectrogram(ae_fft* f, ae_vec* symbols, ae_f32* levels, int use_db);
/* the correlator the crate benchmarks (benches/benches.rs:410-416), in place, `howmany` frames:
 * input.vec_rfft(&mut fft, s).vec_mul(&sig).vec_rifft(&mut fft, s); sig.len == fft.len() */
ae_status ae_correlate(ae_fft* f, ae_vec* inout, ae_vec* sig, int scale_kind, float x, size_t howmany);

/* noise::Awgn::iter() (src/noise.rs:68-84): the next n samples of the stream, to a HOST slice */
ae_status ae_awgn_next_host(ae_awgn* g, ae_cf32* host_out, size_t n);
/* sample-file format of util::file (src/util/file.rs:12-107): raw native-endian cf32 structs back to
 * back, no header.  AE_EARG ("File does not contain an integer number of the requested struct",
 * :19-22) when the size is not a multiple of 8.  Pinned, chunked staging. */
ae_status ae_vec_read_raw(const char* path, ae_vec** out);
ae_status ae_vec_write_raw(ae_vec* v, const char* path);

#ifdef __cplusplus
}
#endif
#endif /* AETHER_B200_H */
