// api.cu — host side of the C ABI declared in include/aether_b200.h: device contexts, buffer
// handles with the lazy VecOps op tape, plans (FFT, FIR, modulation, AWGN, chains) and the
// status-code error model.  No CPU compute path exists here: every data operation launches a
// kernel from elementwise.cu / fft.cu / fir.cu / chain.cu on the context's stream.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <dlfcn.h>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "internal.h"

using namespace ae;

// =================================================================================================
// context / errors
// =================================================================================================
namespace {

struct Alloc {
  void* p = nullptr;
  size_t bytes = 0;
  bool owned = true;
  int refs = 1;
};

struct Ctx {
  int dev = 0;
  cudaStream_t own = nullptr, stream = nullptr;
  int sm_count = 148;
  int* d_err = nullptr;
  std::vector<ae_vec*> pending;               // vecs with a non-empty tape
  std::map<size_t, float2*> tw_cache;         // plain twiddle tables W[k] by length (any-length path)
  std::map<size_t, float2*> ttw_cache;        // per-thread twiddle tables of the power-of-two kernels
  float2* x2tw1024 = nullptr;                 // per-thread twiddle rows of the packed 32x32 transform (chain_x2.cuh)
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
};

constexpr int kMaxDev = 64;
Ctx* g_ctx[kMaxDev] = {nullptr};
std::mutex g_mu;
thread_local int t_dev = -1;
thread_local std::string t_err = "";
std::atomic<uint64_t> g_launches{0};

ae_status fail(ae_status st, const std::string& msg) {
  t_err = msg;
  return st;
}
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      if (e__ == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(AE_EOOM, std::string("out of device memory: ") + #call); } \
      return fail(AE_ECUDA, std::string(cudaGetErrorString(e__)) + " in " + #call);           \
    }                                                                                         \
  } while (0)
#define CKL(n)                                                                                \
  do {                                                                                        \
    if (const char* u__ = ae::take_unsupported_launch()) return fail(AE_EARG, std::string("no kernel for this shape: ") + u__); \
    g_launches += (n);                                                                        \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) return fail(AE_ECUDA, std::string("kernel launch failed: ") + cudaGetErrorString(e__)); \
  } while (0)
#define TRY(expr)                        \
  do {                                   \
    ae_status s__ = (expr);              \
    if (s__ != AE_OK) return s__;        \
  } while (0)

}  // namespace
namespace ae {
namespace {
thread_local const char* t_unsupported = nullptr;
struct ResKey { const void* k; int dev; int threads; size_t smem; bool operator<(const ResKey& o) const { return std::tie(k, dev, threads, smem) < std::tie(o.k, o.dev, o.threads, o.smem); } };
std::map<ResKey, size_t> g_resident;
std::map<std::pair<const void*, int>, size_t> g_optin;   // largest dynamic shared-memory opt-in made per (kernel, device)
std::mutex g_resident_mu;
}  // namespace
void note_unsupported_launch(const char* what) { t_unsupported = what; }
const char* take_unsupported_launch() { const char* w = t_unsupported; t_unsupported = nullptr; return w; }
size_t resident_ctas(const void* kern, int threads, size_t smem) {
  int dev = 0;
  cudaGetDevice(&dev);
  const ResKey key{kern, dev, threads, smem};
  std::lock_guard<std::mutex> lk(g_resident_mu);
  auto it = g_resident.find(key);
  if (it != g_resident.end()) return it->second;
  // the opt-in is one value per kernel: only ever raise it (a kernel is launched with several sizes, e.g. the
  // any-length FFT; lowering it would make an earlier, larger configuration fail with "invalid argument")
  size_t& optin = g_optin[std::make_pair(kern, dev)];
  if (smem > 48 * 1024 && smem > optin) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    optin = smem;
  }
  int per_sm = 1, sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  const size_t r = (size_t)sms * per_sm;
  g_resident[key] = r;
  return r;
}
}  // namespace ae
namespace {

ae_status get_ctx(Ctx** out) {
  if (t_dev < 0) {
    ae_status s = ae_init(0);
    if (s != AE_OK) return s;
  }
  Ctx* c = g_ctx[t_dev];
  cudaError_t e = cudaSetDevice(c->dev);
  if (e != cudaSuccess) return fail(AE_ECUDA, cudaGetErrorString(e));
  *out = c;
  return AE_OK;
}

ae_status dev_alloc(Ctx* c, size_t bytes, void** p) {
  if (bytes == 0) bytes = 256;
  bytes = (bytes + 255) & ~(size_t)255;
  CK(cudaMallocAsync(p, bytes, c->stream));
  return AE_OK;
}
void dev_free(Ctx* c, void* p) {
  if (p) cudaFreeAsync(p, c->stream);
}

Alloc* alloc_new(Ctx* c, size_t bytes, ae_status* st) {
  void* p = nullptr;
  *st = dev_alloc(c, bytes, &p);
  if (*st != AE_OK) return nullptr;
  Alloc* a = new Alloc;
  a->p = p; a->bytes = bytes; a->owned = true; a->refs = 1;
  return a;
}
void alloc_unref(Ctx* c, Alloc* a) {
  if (!a) return;
  if (--a->refs == 0) {
    if (a->owned) dev_free(c, a->p);
    delete a;
  }
}

ae_status get_twiddles(Ctx* c, size_t n, float2** out) {
  auto it = c->tw_cache.find(n);
  if (it != c->tw_cache.end()) { *out = it->second; return AE_OK; }
  std::vector<float2> h(n);
  for (size_t k = 0; k < n; ++k) {
    const double a = -2.0 * M_PI * (double)k / (double)n;   // f64, rounded to f32 (rustfft's accuracy class)
    h[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  void* p = nullptr;
  TRY(dev_alloc(c, n * sizeof(float2), &p));
  CK(cudaMemcpyAsync(p, h.data(), n * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->tw_cache[n] = (float2*)p;
  *out = (float2*)p;
  return AE_OK;
}

ae_status get_thread_twiddles(Ctx* c, size_t n, float2** out) {
  auto it = c->ttw_cache.find(n);
  if (it != c->ttw_cache.end()) { *out = it->second; return AE_OK; }
  std::vector<float2> h;
  fft_thread_twiddles(n, h);
  if (h.empty()) return fail(AE_EARG, "no power-of-two kernel for this length");
  void* p = nullptr;
  TRY(dev_alloc(c, h.size() * sizeof(float2), &p));
  CK(cudaMemcpyAsync(p, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->ttw_cache[n] = (float2*)p;
  *out = (float2*)p;
  return AE_OK;
}

// twiddle rows of the packed one-exchange 1024-point transform (K14b / K4b / correlator), one table per device
ae_status get_x2_twiddles(Ctx* c, float2** out) {
  if (!c->x2tw1024) {
    std::vector<float2> tw, hi, lo;
    const float2 one = make_float2(1.f, 0.f);
    chain_x2_tables(1024, &one, 1, tw, hi, lo);
    void* p = nullptr;
    TRY(dev_alloc(c, tw.size() * sizeof(float2), &p));
    CK(cudaMemcpyAsync(p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->x2tw1024 = (float2*)p;
  }
  *out = c->x2tw1024;
  return AE_OK;
}

}  // namespace

// =================================================================================================
// handle types
// =================================================================================================
struct TapeRec {
  int op;
  float s;
  Alloc* oa;     // operand allocation (ref held) or null
  size_t ooff;   // operand offset in elements
};
struct ae_vec {
  Ctx* c;
  Alloc* a;
  size_t off, len, cap;
  std::vector<TapeRec> tape;
  bool plan_owned;  // scratch view handed out by a plan: ae_vec_free is a no-op
};
struct ae_bits {
  Ctx* c;
  Alloc* a;
  size_t off, len, cap;
};

namespace {

inline float2* vptr(const ae_vec* v) { return reinterpret_cast<float2*>(v->a->p) + v->off; }
inline uint8_t* bptr(const ae_bits* b) { return reinterpret_cast<uint8_t*>(b->a->p) + b->off; }

void drop_pending(ae_vec* v) {
  auto& p = v->c->pending;
  p.erase(std::remove(p.begin(), p.end(), v), p.end());
}

ae_status flush_vec(ae_vec* v) {
  if (v->tape.empty()) return AE_OK;
  TapeParams tp;
  std::memset(&tp, 0, sizeof(tp));
  bool has_mirror = false;
  int n = 0;
  for (const TapeRec& r : v->tape) {
    tp.e[n].op = r.op;
    tp.e[n].s = r.s;
    tp.e[n].operand = r.oa ? reinterpret_cast<const float2*>(r.oa->p) + r.ooff : nullptr;
    has_mirror |= (r.op == OP_MIRROR);
    ++n;
  }
  tp.n_ops = n;
  tp.load_self = !(v->tape[0].op == OP_ZERO || v->tape[0].op == OP_CLONE);
  launch_vecops(vptr(v), v->len, tp, has_mirror, v->c->sm_count, v->c->stream);
  for (TapeRec& r : v->tape) alloc_unref(v->c, r.oa);
  v->tape.clear();
  drop_pending(v);
  if (v->len) CKL(1);
  return AE_OK;
}

// Two allocations alias when their device address ranges overlap: an Alloc is unique for library
// memory, but ae_vec_wrap / ae_bits_wrap can wrap the same foreign pointer more than once.
bool same_mem(const Alloc* x, const Alloc* y) {
  if (x == y) return true;
  if (!x || !y || (x->owned && y->owned)) return false;
  const char *xb = (const char*)x->p, *yb = (const char*)y->p;
  return xb < yb + y->bytes && yb < xb + x->bytes;
}
bool tape_reads(const ae_vec* p, const Alloc* a) {
  for (const TapeRec& r : p->tape)
    if (r.oa && same_mem(r.oa, a)) return true;
  return false;
}
// make the memory behind `v` current (its own tape and any alias of the same allocation)
ae_status before_read(ae_vec* v) {
  std::vector<ae_vec*> snap = v->c->pending;
  for (ae_vec* p : snap)
    if (same_mem(p->a, v->a)) TRY(flush_vec(p));
  return AE_OK;
}
// additionally run every tape that still wants to read the old contents of `v`
ae_status before_write(ae_vec* v) {
  std::vector<ae_vec*> snap = v->c->pending;
  for (ae_vec* p : snap)
    if (same_mem(p->a, v->a) || tape_reads(p, v->a)) TRY(flush_vec(p));
  return AE_OK;
}

ae_status record(ae_vec* v, int op, ae_vec* other, float s) {
  if (!v) return fail(AE_EARG, "null vector handle");
  const bool binary = (op == OP_MUL || op == OP_DIV || op == OP_ADD || op == OP_SUB || op == OP_CLONE);
  if (binary) {
    if (!other) return fail(AE_EARG, "null operand handle");
    if (other->len != v->len) return fail(AE_ELEN, "Vectors must have same length");  // src/vecops.rs:100-104
    if (other->c != v->c) return fail(AE_EARG, "operand lives on another device");
  }
  std::vector<ae_vec*> snap = v->c->pending;
  for (ae_vec* p : snap) {
    if (p == v) continue;
    if (same_mem(p->a, v->a) || tape_reads(p, v->a) || (binary && same_mem(p->a, other->a))) TRY(flush_vec(p));
  }
  if (binary && same_mem(other->a, v->a)) TRY(flush_vec(v));  // operand aliases self: snapshot current values
  if ((int)v->tape.size() >= kMaxTape) TRY(flush_vec(v));
  if (op == OP_ZERO || op == OP_CLONE) {  // everything recorded before is dead
    for (TapeRec& r : v->tape) alloc_unref(v->c, r.oa);
    v->tape.clear();
  }
  TapeRec r;
  r.op = op; r.s = s; r.oa = nullptr; r.ooff = 0;
  if (binary) { r.oa = other->a; r.ooff = other->off; other->a->refs++; }
  if (v->tape.empty()) v->c->pending.push_back(v);
  v->tape.push_back(r);
  return AE_OK;
}

ae_status vec_reserve(ae_vec* v, size_t cap) {
  if (cap <= v->cap) return AE_OK;
  if (!v->a->owned || v->plan_owned) return fail(AE_EARG, "cannot grow a borrowed or plan-owned vector");
  TRY(before_write(v));
  size_t ncap = std::max(cap, v->cap * 2);
  ae_status st;
  Alloc* na = alloc_new(v->c, ncap * sizeof(float2), &st);
  if (!na) return st;
  if (v->len) CK(cudaMemcpyAsync(na->p, vptr(v), v->len * sizeof(float2), cudaMemcpyDeviceToDevice, v->c->stream));
  alloc_unref(v->c, v->a);
  v->a = na; v->off = 0; v->cap = ncap;
  return AE_OK;
}
ae_status bits_reserve(ae_bits* b, size_t cap) {
  if (cap <= b->cap) return AE_OK;
  if (!b->a->owned) return fail(AE_EARG, "cannot grow a borrowed bit vector");
  size_t ncap = std::max(cap, b->cap * 2);
  ae_status st;
  Alloc* na = alloc_new(b->c, ncap, &st);
  if (!na) return st;
  if (b->len) CK(cudaMemcpyAsync(na->p, bptr(b), b->len, cudaMemcpyDeviceToDevice, b->c->stream));
  alloc_unref(b->c, b->a);
  b->a = na; b->off = 0; b->cap = ncap;
  return AE_OK;
}

float scale_factor(int kind, size_t n, float x) {  // src/fft.rs:22-37
  switch (kind) {
    case AE_SCALE_SN: return 1.0f / sqrtf((float)n);
    case AE_SCALE_N: return 1.0f / (float)n;
    case AE_SCALE_X: return x;
    default: return 1.0f;
  }
}

}  // namespace

// =================================================================================================
// runtime
// =================================================================================================
extern "C" {

const char* ae_version(void) { return "aether_b200 0.1.0 (sm_100a)"; }

ae_status ae_device_count(int* n) {
  if (!n) return fail(AE_EARG, "null");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) { *n = 0; cudaGetLastError(); return fail(AE_ECUDA, cudaGetErrorString(e)); }
  return AE_OK;
}

ae_status ae_init(int device) {
  if (device < 0 || device >= kMaxDev) return fail(AE_EARG, "bad device index");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(AE_ECUDA, "no CUDA device available: aether_b200 has no CPU fallback");
  }
  if (device >= n) return fail(AE_EARG, "device index out of range");
  CK(cudaSetDevice(device));
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx[device]) {
    Ctx* c = new Ctx;
    c->dev = device;
    // a failure half-way releases what was created (the mutex guard unlocks on every return)
    auto build = [&]() -> ae_status {
      CK(cudaStreamCreateWithFlags(&c->own, cudaStreamNonBlocking));
      c->stream = c->own;
      for (int i = 0; i < 3; ++i) CK(cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
      cudaDeviceProp prop;
      CK(cudaGetDeviceProperties(&prop, device));
      c->sm_count = prop.multiProcessorCount;
      if (prop.major != 10)
        return fail(AE_ECUDA, "aether_b200 kernels are built for sm_100a (B200) only; found compute capability " +
                                  std::to_string(prop.major) + "." + std::to_string(prop.minor));
      CK(cudaMalloc((void**)&c->d_err, sizeof(int)));
      CK(cudaMemset(c->d_err, 0, sizeof(int)));
      return AE_OK;
    };
    const ae_status st = build();
    if (st != AE_OK) {
      if (c->own) cudaStreamDestroy(c->own);
      for (int i = 0; i < 3; ++i) if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
      if (c->d_err) cudaFree(c->d_err);
      delete c;
      return st;
    }
    // keep freed blocks in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    g_ctx[device] = c;
  }
  t_dev = device;
  return AE_OK;
}

ae_status ae_set_stream(void* cuda_stream) {
  Ctx* c;
  TRY(get_ctx(&c));
  cudaStream_t ns = cuda_stream ? (cudaStream_t)cuda_stream : c->own;
  if (ns != c->stream) {
    // order the new stream after everything already issued on the old one
    cudaEvent_t ev;
    CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CK(cudaEventRecord(ev, c->stream));
    CK(cudaStreamWaitEvent(ns, ev, 0));
    CK(cudaEventDestroy(ev));
    c->stream = ns;
  }
  return AE_OK;
}
void* ae_get_stream(void) {
  Ctx* c;
  if (get_ctx(&c) != AE_OK) return nullptr;
  return (void*)c->stream;
}

}  // extern "C"
// synchronise ONE context (the handle's, which need not be the calling thread's current device) and
// report its deferred device-side error flag
static ae_status sync_ctx(Ctx* c) {
  cudaSetDevice(c->dev);
  int flag = 0;
  CK(cudaMemcpyAsync(&flag, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (flag) {
    CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    if (flag & DEVERR_MOD_INDEX) return fail(AE_EIDX, "index out of bounds: the len is 2 or 4 but a bit pattern selected a larger table index");
    return fail(AE_ECUDA, "device-side error flag set");
  }
  return AE_OK;
}
extern "C" {
ae_status ae_sync(void) {
  Ctx* c;
  TRY(get_ctx(&c));
  return sync_ctx(c);
}
const char* ae_last_error_string(void) { return t_err.c_str(); }
ae_status ae_sm_count(int* n) {
  Ctx* c;
  TRY(get_ctx(&c));
  *n = c->sm_count;
  return AE_OK;
}
uint64_t ae_launch_count(void) { return g_launches.load(); }

ae_status ae_host_alloc(size_t bytes, void** p) {
  Ctx* c;
  TRY(get_ctx(&c));
  CK(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return AE_OK;
}
ae_status ae_host_free(void* p) {
  if (p) CK(cudaFreeHost(p));
  return AE_OK;
}

// =================================================================================================
// ae_vec
// =================================================================================================
ae_status ae_vec_alloc(size_t len, size_t capacity, ae_vec** out) {
  if (!out) return fail(AE_EARG, "null out");
  Ctx* c;
  TRY(get_ctx(&c));
  if (capacity < len) capacity = len;
  ae_status st;
  Alloc* a = alloc_new(c, capacity * sizeof(float2), &st);
  if (!a) return st;
  if (len) CK(cudaMemsetAsync(a->p, 0, len * sizeof(float2), c->stream));
  ae_vec* v = new ae_vec;
  v->c = c; v->a = a; v->off = 0; v->len = len; v->cap = capacity; v->plan_owned = false;
  *out = v;
  return AE_OK;
}
ae_status ae_vec_wrap(void* device_ptr, size_t len, ae_vec** out) {
  if (!out || (!device_ptr && len)) return fail(AE_EARG, "null");
  if (((uintptr_t)device_ptr % 8) != 0) return fail(AE_EARG, "cf32 buffers must be 8-byte aligned");
  Ctx* c;
  TRY(get_ctx(&c));
  Alloc* a = new Alloc;
  a->p = device_ptr; a->bytes = len * sizeof(float2); a->owned = false; a->refs = 1;
  ae_vec* v = new ae_vec;
  v->c = c; v->a = a; v->off = 0; v->len = len; v->cap = len; v->plan_owned = false;
  *out = v;
  return AE_OK;
}
ae_status ae_vec_view(ae_vec* parent, size_t offset, size_t len, ae_vec** out) {
  if (!parent || !out) return fail(AE_EARG, "null");
  if (offset > parent->len || len > parent->len - offset) return fail(AE_EIDX, "range end index out of range for slice");
  ae_vec* v = new ae_vec;
  v->c = parent->c; v->a = parent->a; parent->a->refs++;
  v->off = parent->off + offset; v->len = len; v->cap = len; v->plan_owned = false;
  *out = v;
  return AE_OK;
}
ae_status ae_vec_free(ae_vec* v) {
  if (!v || v->plan_owned) return AE_OK;
  cudaSetDevice(v->c->dev);
  ae_status st = AE_OK;
  if (!v->a->owned) {
    // borrowed memory (ae_vec_wrap): its owner observes it after the handle is gone, and may free it.
    // Run the handle's own tape AND every pending tape that still reads this memory.
    std::vector<ae_vec*> snap = v->c->pending;
    for (ae_vec* p : snap)
      if (p == v || same_mem(p->a, v->a) || tape_reads(p, v->a)) { ae_status s2 = flush_vec(p); if (st == AE_OK) st = s2; }
  } else if (v->a->refs > 1) st = flush_vec(v);  // someone else can still observe the memory
  else {
    for (TapeRec& r : v->tape) alloc_unref(v->c, r.oa);
    v->tape.clear();
    drop_pending(v);
  }
  // tapes that read an owned allocation keep it alive through their own reference
  alloc_unref(v->c, v->a);
  delete v;
  return st;
}
size_t ae_vec_len(const ae_vec* v) { return v ? v->len : 0; }
size_t ae_vec_capacity(const ae_vec* v) { return v ? v->cap : 0; }
ae_status ae_vec_set_len(ae_vec* v, size_t len) {
  if (!v) return fail(AE_EARG, "null");
  if (len > v->cap) return fail(AE_EARG, "len exceeds capacity");
  TRY(flush_vec(v));
  v->len = len;
  return AE_OK;
}
ae_status ae_vec_reserve(ae_vec* v, size_t capacity) {
  if (!v) return fail(AE_EARG, "null");
  cudaSetDevice(v->c->dev);
  return vec_reserve(v, capacity);
}
ae_status ae_vec_device_ptr(ae_vec* v, void** ptr) {
  if (!v || !ptr) return fail(AE_EARG, "null");
  cudaSetDevice(v->c->dev);
  TRY(before_write(v));
  *ptr = vptr(v);
  return AE_OK;
}
ae_status ae_vec_upload(ae_vec* v, const ae_cf32* host, size_t n) {
  if (!v || (!host && n)) return fail(AE_EARG, "null");
  if (n != v->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(v->c->dev);
  TRY(before_write(v));
  if (n) {
    CK(cudaMemcpyAsync(vptr(v), host, n * sizeof(float2), cudaMemcpyHostToDevice, v->c->stream));
    CK(cudaStreamSynchronize(v->c->stream));
  }
  return AE_OK;
}
ae_status ae_vec_upload_async(ae_vec* v, const ae_cf32* host, size_t n) {
  if (!v || (!host && n)) return fail(AE_EARG, "null");
  if (n != v->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(v->c->dev);
  TRY(before_write(v));
  if (n) CK(cudaMemcpyAsync(vptr(v), host, n * sizeof(float2), cudaMemcpyHostToDevice, v->c->stream));
  return AE_OK;
}
ae_status ae_vec_download_async(ae_vec* v, ae_cf32* host, size_t n) {
  if (!v || (!host && n)) return fail(AE_EARG, "null");
  if (n != v->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(v->c->dev);
  TRY(before_read(v));
  if (n) CK(cudaMemcpyAsync(host, vptr(v), n * sizeof(float2), cudaMemcpyDeviceToHost, v->c->stream));
  return AE_OK;
}
ae_status ae_vec_download(ae_vec* v, ae_cf32* host, size_t n) {
  if (!v || (!host && n)) return fail(AE_EARG, "null");
  if (n != v->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(v->c->dev);
  TRY(before_read(v));
  if (n) CK(cudaMemcpyAsync(host, vptr(v), n * sizeof(float2), cudaMemcpyDeviceToHost, v->c->stream));
  return sync_ctx(v->c);
}

ae_status ae_vec_scale(ae_vec* v, float s) { return record(v, OP_SCALE, nullptr, s); }
ae_status ae_vec_mul(ae_vec* v, ae_vec* o) { return record(v, OP_MUL, o, 0.f); }
ae_status ae_vec_div(ae_vec* v, ae_vec* o) { return record(v, OP_DIV, o, 0.f); }
ae_status ae_vec_conj(ae_vec* v) { return record(v, OP_CONJ, nullptr, 0.f); }
ae_status ae_vec_add(ae_vec* v, ae_vec* o) { return record(v, OP_ADD, o, 0.f); }
ae_status ae_vec_sub(ae_vec* v, ae_vec* o) { return record(v, OP_SUB, o, 0.f); }
ae_status ae_vec_mirror(ae_vec* v) { return record(v, OP_MIRROR, nullptr, 0.f); }
ae_status ae_vec_clone(ae_vec* v, ae_vec* o) { return record(v, OP_CLONE, o, 0.f); }
ae_status ae_vec_zero(ae_vec* v) { return record(v, OP_ZERO, nullptr, 0.f); }
ae_status ae_vec_flush(ae_vec* v) {
  if (!v) return fail(AE_EARG, "null");
  cudaSetDevice(v->c->dev);
  return flush_vec(v);
}
size_t ae_vec_pending_ops(const ae_vec* v) { return v ? v->tape.size() : 0; }

ae_status ae_vec_mutate(ae_vec* v, void (*f)(ae_cf32*, void*), void* user) {
  if (!v || !f) return fail(AE_EARG, "null");
  // arbitrary host closure in element order (src/vecops.rs:179-182): documented slow path
  std::vector<ae_cf32> h(v->len);
  TRY(ae_vec_download(v, h.data(), v->len));
  for (size_t i = 0; i < v->len; ++i) f(&h[i], user);
  return ae_vec_upload(v, h.data(), v->len);
}

ae_status ae_scale_factor(int kind, size_t n, float x, float* s) {
  if (!s || kind < 0 || kind > 3) return fail(AE_EARG, "bad scale kind");
  *s = scale_factor(kind, n, x);
  return AE_OK;
}
ae_status ae_vec_scale_kind(ae_vec* v, int kind, float x) {
  if (!v || kind < 0 || kind > 3) return fail(AE_EARG, "bad scale kind");
  if (kind == AE_SCALE_NONE) return AE_OK;  // src/fft.rs:24
  return record(v, OP_SCALE, nullptr, scale_factor(kind, v->len, x));
}

// =================================================================================================
// ae_bits
// =================================================================================================
ae_status ae_bits_alloc(size_t len, size_t capacity, ae_bits** out) {
  if (!out) return fail(AE_EARG, "null out");
  Ctx* c;
  TRY(get_ctx(&c));
  if (capacity < len) capacity = len;
  ae_status st;
  Alloc* a = alloc_new(c, capacity, &st);
  if (!a) return st;
  if (len) CK(cudaMemsetAsync(a->p, 0, len, c->stream));
  ae_bits* b = new ae_bits;
  b->c = c; b->a = a; b->off = 0; b->len = len; b->cap = capacity;
  *out = b;
  return AE_OK;
}
ae_status ae_bits_wrap(void* device_ptr, size_t len, ae_bits** out) {
  if (!out || (!device_ptr && len)) return fail(AE_EARG, "null");
  Ctx* c;
  TRY(get_ctx(&c));
  Alloc* a = new Alloc;
  a->p = device_ptr; a->bytes = len; a->owned = false; a->refs = 1;
  ae_bits* b = new ae_bits;
  b->c = c; b->a = a; b->off = 0; b->len = len; b->cap = len;
  *out = b;
  return AE_OK;
}
ae_status ae_bits_free(ae_bits* b) {
  if (!b) return AE_OK;
  cudaSetDevice(b->c->dev);
  alloc_unref(b->c, b->a);
  delete b;
  return AE_OK;
}
size_t ae_bits_len(const ae_bits* b) { return b ? b->len : 0; }
size_t ae_bits_capacity(const ae_bits* b) { return b ? b->cap : 0; }
ae_status ae_bits_set_len(ae_bits* b, size_t len) {
  if (!b) return fail(AE_EARG, "null");
  if (len > b->cap) return fail(AE_EARG, "len exceeds capacity");
  b->len = len;
  return AE_OK;
}
ae_status ae_bits_device_ptr(ae_bits* b, void** ptr) {
  if (!b || !ptr) return fail(AE_EARG, "null");
  *ptr = bptr(b);
  return AE_OK;
}
ae_status ae_bits_upload(ae_bits* b, const uint8_t* host, size_t n) {
  if (!b || (!host && n)) return fail(AE_EARG, "null");
  if (n != b->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(b->c->dev);
  if (n) {
    CK(cudaMemcpyAsync(bptr(b), host, n, cudaMemcpyHostToDevice, b->c->stream));
    CK(cudaStreamSynchronize(b->c->stream));
  }
  return AE_OK;
}
ae_status ae_bits_upload_async(ae_bits* b, const uint8_t* host, size_t n) {
  if (!b || (!host && n)) return fail(AE_EARG, "null");
  if (n != b->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(b->c->dev);
  if (n) CK(cudaMemcpyAsync(bptr(b), host, n, cudaMemcpyHostToDevice, b->c->stream));
  return AE_OK;
}
ae_status ae_bits_download_async(ae_bits* b, uint8_t* host, size_t n) {
  if (!b || (!host && n)) return fail(AE_EARG, "null");
  if (n != b->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(b->c->dev);
  if (n) CK(cudaMemcpyAsync(host, bptr(b), n, cudaMemcpyDeviceToHost, b->c->stream));
  return AE_OK;
}
ae_status ae_bits_download(ae_bits* b, uint8_t* host, size_t n) {
  if (!b || (!host && n)) return fail(AE_EARG, "null");
  if (n != b->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(b->c->dev);
  if (n) CK(cudaMemcpyAsync(host, bptr(b), n, cudaMemcpyDeviceToHost, b->c->stream));
  return sync_ctx(b->c);
}

}  // extern "C"

// =================================================================================================
// FFT
// =================================================================================================
struct ae_fft {
  Ctx* c;
  size_t len;
  int compat;
  float2* tw;
  bool pow2;
  bool big = false;                       // four-step path (2^15..2^24)
  float2 *tw1 = nullptr, *tw2 = nullptr;  // per-thread tables of the two factor lengths
  float2 *wlo = nullptr, *whi = nullptr;  // two-level W_len table (owned)
  std::vector<uint32_t> radices;
  // Bluestein path (non-power-of-two lengths > 6144 or with a prime factor > 61)
  bool blue = false;
  unsigned blue_log2m = 0;
  ae_fft* blue_child = nullptr;            // power-of-two plan of length M = 2^blue_log2m
  float2 *blue_w = nullptr, *blue_b = nullptr;   // chirp w[n] (len entries), FFT_M of the wrapped conj chirp (M entries)
  float2* scratch;
  size_t scratch_elems;
  Alloc* tmp;       // tfwd/tbwd result buffer
  size_t tmp_elems;
  ae_vec tmpview;
};

namespace {

ae_status fft_run(ae_fft* f, int dir, const float2* in, float2* out, int scale_kind, float x, size_t howmany) {
  Ctx* c = f->c;
  // reference compat: Cfft "fwd" = rustfft inverse = exp(+) (src/fft.rs:148)
  const bool fwd_is_inverse = (f->compat == AE_COMPAT_REFERENCE);
  const bool inverse = (dir == AE_FFT_FWD) ? fwd_is_inverse : !fwd_is_inverse;
  const bool do_scale = scale_kind != AE_SCALE_NONE;
  const float s = scale_factor(scale_kind, f->len, x);  // N = the slice length passed (one frame)
  if (howmany == 0) return AE_OK;
  if (f->pow2) {
    launch_fft_pow2(in, out, f->len, howmany, f->tw, inverse, do_scale, s, c->stream);
    CKL(1);
  } else if (f->big) {
    const size_t need = f->len * howmany;
    if (f->scratch_elems < need) {
      dev_free(c, f->scratch);
      f->scratch = nullptr; f->scratch_elems = 0;
      void* p;
      TRY(dev_alloc(c, need * sizeof(float2), &p));
      f->scratch = (float2*)p; f->scratch_elems = need;
    }
    launch_fft_big(in, out, f->scratch, f->len, howmany, f->tw1, f->tw2, f->wlo, f->whi, inverse, do_scale, s, c->stream);
    CKL(2 * ((howmany + 32767) / 32768));
  } else if (f->blue) {
    const size_t m = (size_t)1 << f->blue_log2m;
    const size_t need = m * howmany;
    if (f->scratch_elems < need) {
      dev_free(c, f->scratch);
      f->scratch = nullptr; f->scratch_elems = 0;
      void* p;
      TRY(dev_alloc(c, need * sizeof(float2), &p));
      f->scratch = (float2*)p; f->scratch_elems = need;
    }
    launch_bluestein_pre(in, f->scratch, f->blue_w, f->len, f->blue_log2m, howmany, inverse, c->sm_count, c->stream);
    CKL(1);
    TRY(fft_run(f->blue_child, AE_FFT_FWD, f->scratch, f->scratch, AE_SCALE_NONE, 1.0f, howmany));
    launch_bluestein_mul(f->scratch, f->blue_b, f->blue_log2m, howmany, c->sm_count, c->stream);
    CKL(1);
    TRY(fft_run(f->blue_child, AE_FFT_BWD, f->scratch, f->scratch, AE_SCALE_NONE, 1.0f, howmany));
    launch_bluestein_post(f->scratch, out, f->blue_w, f->len, f->blue_log2m, howmany, inverse, do_scale, s, c->stream);
    CKL((howmany + 32767) / 32768);
  } else {
    const size_t need = 2 * f->len * howmany;
    if (f->scratch_elems < need) {
      dev_free(c, f->scratch);
      f->scratch = nullptr; f->scratch_elems = 0;
      void* p;
      TRY(dev_alloc(c, need * sizeof(float2), &p));
      f->scratch = (float2*)p; f->scratch_elems = need;
    }
    launch_fft_generic(in, out, f->scratch, f->len, howmany, f->tw, f->radices.data(), (int)f->radices.size(), inverse,
                       do_scale, s, c->stream);
    CKL(f->len <= 6144 ? 1 : std::max<size_t>(1, f->radices.size()) * ((howmany + 32767) / 32768));
  }
  return AE_OK;
}

}  // namespace

extern "C" {

ae_status ae_fft_create(size_t len, ae_fft** out) {
  if (!out) return fail(AE_EARG, "null out");
  if (len == 0) return fail(AE_EARG, "FFT length must be >= 1");
  if (len > (1ull << 31)) return fail(AE_EARG, "FFT length too large");
  Ctx* c;
  TRY(get_ctx(&c));
  ae_fft* f = new ae_fft;
  f->c = c; f->len = len; f->compat = AE_COMPAT_REFERENCE;
  f->pow2 = fft_pow2_supported(len);
  f->scratch = nullptr; f->scratch_elems = 0; f->tmp = nullptr; f->tmp_elems = 0;
  f->big = !f->pow2 && fft_big_supported(len);
  ae_status st = AE_OK;
  if (f->big) {
    size_t n1, n2;
    fft_big_split(len, &n1, &n2);
    st = get_thread_twiddles(c, n1, &f->tw1);
    if (st == AE_OK) st = get_thread_twiddles(c, n2, &f->tw2);
    if (st == AE_OK) {
      std::vector<float2> lo(4096), hi(len / 4096);
      for (size_t e = 0; e < 4096; ++e) {
        const double a = -2.0 * M_PI * (double)e / (double)len;
        lo[e] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
      for (size_t e = 0; e < hi.size(); ++e) {
        const double a = -2.0 * M_PI * (double)(e * 4096) / (double)len;
        hi[e] = make_float2((float)std::cos(a), (float)std::sin(a));
      }
      void* p;
      st = dev_alloc(c, lo.size() * sizeof(float2), &p);
      if (st == AE_OK) {
        f->wlo = (float2*)p;
        cudaMemcpyAsync(f->wlo, lo.data(), lo.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
        st = dev_alloc(c, hi.size() * sizeof(float2), &p);
      }
      if (st == AE_OK) {
        f->whi = (float2*)p;
        cudaMemcpyAsync(f->whi, hi.data(), hi.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);
      }
    }
    f->tw = nullptr;
  } else {
    st = f->pow2 ? get_thread_twiddles(c, len, &f->tw) : get_twiddles(c, len, &f->tw);
  }
  if (st != AE_OK) { delete f; return st; }
  if (!f->pow2 && !f->big) {
    size_t m = len;
    while (m > 1) {
      size_t p = 0;
      for (size_t q : {4, 2, 3, 5, 7}) if (m % q == 0) { p = q; break; }
      if (!p) {
        p = 11;
        while (m % p) { p += 2; if (p * p > m) { p = m; break; } }
      }
      f->radices.push_back((uint32_t)p);
      m /= p;
    }
  }
  f->tmpview.c = c; f->tmpview.a = nullptr; f->tmpview.off = 0; f->tmpview.len = 0; f->tmpview.cap = 0;
  f->tmpview.plan_owned = true;
  if (!f->pow2 && !f->big) {
    uint32_t pmax = 1;
    for (uint32_t r : f->radices) pmax = std::max(pmax, r);
    if ((len & (len - 1)) != 0 && (len > 6144 || pmax > 61)) {   // (a power of two beyond the four-step range keeps the per-factor path)
      // Bluestein: M = the power of two >= 2*len - 1; chirp exponents n^2 mod 2N are exact integers
      unsigned lg = 1;
      while (((size_t)1 << lg) < 2 * len - 1) ++lg;
      const size_t m = (size_t)1 << lg;
      ae_status st2 = ae_fft_create(m, &f->blue_child);
      if (st2 == AE_OK) st2 = ae_fft_set_compat(f->blue_child, AE_COMPAT_CORRECTED);   // FWD = exp(-)
      std::vector<float2> w(len), b(m, make_float2(0.0f, 0.0f));
      for (size_t n = 0; n < len; ++n) {
        const unsigned long long e = ((unsigned long long)n * n) % (2ull * len);
        const double a = -M_PI * (double)e / (double)len;
        w[n] = make_float2((float)std::cos(a), (float)std::sin(a));
        const float2 cw = make_float2(w[n].x, -w[n].y);
        b[n] = cw;
        if (n) b[m - n] = cw;
      }
      void* p = nullptr;
      if (st2 == AE_OK) st2 = dev_alloc(c, len * sizeof(float2), &p);
      if (st2 == AE_OK) {
        f->blue_w = (float2*)p;
        cudaMemcpyAsync(f->blue_w, w.data(), len * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
        st2 = dev_alloc(c, m * sizeof(float2), &p);
      }
      if (st2 == AE_OK) {
        f->blue_b = (float2*)p;
        cudaMemcpyAsync(f->blue_b, b.data(), m * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
        st2 = fft_run(f->blue_child, AE_FFT_FWD, f->blue_b, f->blue_b, AE_SCALE_NONE, 1.0f, 1);
      }
      if (st2 == AE_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st2 = fail(AE_ECUDA, "Bluestein plan setup failed");
      if (st2 != AE_OK) { ae_fft_destroy(f); return st2; }
      f->blue = true;
      f->blue_log2m = lg;
    }
  }
  *out = f;
  return AE_OK;
}
ae_status ae_fft_destroy(ae_fft* f) {
  if (!f) return AE_OK;
  cudaSetDevice(f->c->dev);
  dev_free(f->c, f->scratch);
  dev_free(f->c, f->wlo);
  dev_free(f->c, f->whi);
  dev_free(f->c, f->blue_w); dev_free(f->c, f->blue_b);
  ae_fft_destroy(f->blue_child);
  if (f->tmp) { flush_vec(&f->tmpview); alloc_unref(f->c, f->tmp); }
  delete f;
  return AE_OK;
}
size_t ae_fft_len(const ae_fft* f) { return f ? f->len : 0; }
ae_status ae_fft_set_compat(ae_fft* f, int compat) {
  if (!f || (compat != AE_COMPAT_REFERENCE && compat != AE_COMPAT_CORRECTED)) return fail(AE_EARG, "bad compat");
  f->compat = compat;
  return AE_OK;
}

ae_status ae_fft_exec(ae_fft* f, int dir, ae_vec* in, ae_vec* out, int scale_kind, float x, size_t howmany) {
  if (!f || !in) return fail(AE_EARG, "null");
  if (dir != AE_FFT_FWD && dir != AE_FFT_BWD) return fail(AE_EARG, "bad direction");
  if (scale_kind < 0 || scale_kind > 3) return fail(AE_EARG, "bad scale kind");
  if (in->len != f->len * howmany) return fail(AE_ELEN, "Input and FFT must be the same length");  // src/fft.rs:163-167
  if (out && out->len != in->len) return fail(AE_ELEN, "Input and FFT must be the same length");
  cudaSetDevice(f->c->dev);
  if (!out || out == in || (out->a == in->a && out->off == in->off)) {
    TRY(before_write(in));
    return fft_run(f, dir, vptr(in), vptr(in), scale_kind, x, howmany);
  }
  if (out->a == in->a) return fail(AE_EARG, "fwd/bwd: output overlaps input");
  TRY(before_read(in));
  TRY(before_write(out));
  return fft_run(f, dir, vptr(in), vptr(out), scale_kind, x, howmany);
}

ae_status ae_fft_exec_tmp(ae_fft* f, int dir, ae_vec* in, int scale_kind, float x, size_t howmany, ae_vec** view) {
  if (!f || !in || !view) return fail(AE_EARG, "null");
  if (dir != AE_FFT_FWD && dir != AE_FFT_BWD) return fail(AE_EARG, "bad direction");
  if (in->len != f->len * howmany) return fail(AE_ELEN, "Input and FFT must be the same length");  // src/fft.rs:207-211
  cudaSetDevice(f->c->dev);
  const size_t need = in->len;
  if (f->tmp) TRY(before_write(&f->tmpview));
  if (f->tmp_elems < need || !f->tmp) {
    if (f->tmp) alloc_unref(f->c, f->tmp);
    ae_status st;
    f->tmp = alloc_new(f->c, std::max<size_t>(need, 1) * sizeof(float2), &st);
    if (!f->tmp) { f->tmp_elems = 0; return st; }
    f->tmp_elems = need;
  }
  f->tmpview.a = f->tmp; f->tmpview.off = 0; f->tmpview.len = need; f->tmpview.cap = need;
  if (in->a == f->tmp) return fail(AE_EARG, "tfwd/tbwd: input is the plan's own scratch");
  TRY(before_read(in));
  TRY(fft_run(f, dir, vptr(in), vptr(&f->tmpview), scale_kind, x, howmany));
  *view = &f->tmpview;
  return AE_OK;
}

static ae_status vec_fft_onfly(ae_vec* v, int dir, int scale_kind, float x, int compat) {
  if (!v) return fail(AE_EARG, "null");
  if (v->len == 0) return AE_OK;
  ae_fft* f;
  TRY(ae_fft_create(v->len, &f));  // Cfft::with_len(self.len()) per call (src/vecops.rs:302)
  ae_fft_set_compat(f, compat);
  ae_status st = ae_fft_exec(f, dir, v, nullptr, scale_kind, x, 1);
  ae_fft_destroy(f);
  return st;
}
ae_status ae_vec_fft(ae_vec* v, int scale_kind, float x, int compat) { return vec_fft_onfly(v, AE_FFT_FWD, scale_kind, x, compat); }
ae_status ae_vec_ifft(ae_vec* v, int scale_kind, float x, int compat) { return vec_fft_onfly(v, AE_FFT_BWD, scale_kind, x, compat); }

}  // extern "C"

// =================================================================================================
// FIR
// =================================================================================================
struct ae_fir {
  Ctx* c;
  size_t ntaps;
  int tp;         // taps rounded up to a multiple of 8
  int mode;       // resolved: DIRECT or OVERLAP_SAVE
  float2* d_taps; // tp entries, zero padded
  float2* d_hist; // tp-1 entries, newest last
  float2* d_hist2;
  size_t nfft;
  float2* d_H;
  float2* d_tw;
  float2* d_x2tw;  // K4b (1024-point blocks): per-thread twiddle rows of the packed transform, else null
  std::vector<float2> h_taps;   // host copy, tp entries: the direct kernel takes short filters as kernel parameters
};

extern "C" {

ae_status ae_fir_create(const ae_cf32* taps_host, size_t ntaps, int mode, ae_fir** out) {
  if (!out || !taps_host) return fail(AE_EARG, "null");
  if (ntaps == 0) return fail(AE_EARG, "FIR needs at least one tap");
  if (mode < AE_FIR_AUTO || mode > AE_FIR_OVERLAP_SAVE) return fail(AE_EARG, "bad FIR mode");
  Ctx* c;
  TRY(get_ctx(&c));
  size_t nfft = 1024;
  while (nfft < 4 * ntaps) nfft <<= 1;
  if (const char* e = getenv("AE_FIR_NFFT")) {   // developer override of the overlap-save block length
    const size_t v = (size_t)atoll(e);
    if (v >= 2 * ntaps) nfft = v;
  }
  const bool os_ok = fir_os_supported(nfft) && ntaps <= nfft / 2;
  const bool direct_ok = ntaps <= 4096;
  // measured (tools/fir_quick.py sweep): overlap-save runs at 240-245 Gsamples/s for every tap count it
  // supports, the direct form at 150 (2-8 taps) falling to 75 (64 taps), so AUTO means overlap-save
  // whenever a block length exists; the direct form stays for explicit requests (f32 MACs in tap order)
  if (mode == AE_FIR_AUTO) mode = os_ok ? AE_FIR_OVERLAP_SAVE : AE_FIR_DIRECT;
  if (mode == AE_FIR_DIRECT && !direct_ok) return fail(AE_EARG, "direct-form FIR supports at most 4096 taps");
  if (mode == AE_FIR_OVERLAP_SAVE && !os_ok) return fail(AE_EARG, "overlap-save FIR supports at most 4096 taps");
  ae_fir* f = new ae_fir;
  f->c = c; f->ntaps = ntaps; f->tp = (int)(((ntaps + 7) / 8) * 8); f->mode = mode;
  f->d_taps = f->d_hist = f->d_hist2 = f->d_H = f->d_tw = f->d_x2tw = nullptr; f->nfft = nfft;
  auto build = [&]() -> ae_status {
    void* p;
    std::vector<float2> h(f->tp, make_float2(0.f, 0.f));
    for (size_t i = 0; i < ntaps; ++i) h[i] = make_float2(taps_host[i].re, taps_host[i].im);
    f->h_taps = h;
    TRY(dev_alloc(c, f->tp * sizeof(float2), &p)); f->d_taps = (float2*)p;
    CK(cudaMemcpyAsync(f->d_taps, h.data(), f->tp * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
    TRY(dev_alloc(c, f->tp * sizeof(float2), &p)); f->d_hist = (float2*)p;
    TRY(dev_alloc(c, f->tp * sizeof(float2), &p)); f->d_hist2 = (float2*)p;
    CK(cudaMemsetAsync(f->d_hist, 0, f->tp * sizeof(float2), c->stream));
    CK(cudaMemsetAsync(f->d_hist2, 0, f->tp * sizeof(float2), c->stream));
    if (mode == AE_FIR_OVERLAP_SAVE) {
      TRY(get_thread_twiddles(c, nfft, &f->d_tw));
      std::vector<float2> hp(nfft, make_float2(0.f, 0.f));
      for (size_t i = 0; i < ntaps; ++i) hp[i] = h[i];
      TRY(dev_alloc(c, nfft * sizeof(float2), &p)); f->d_H = (float2*)p;
      CK(cudaMemcpyAsync(f->d_H, hp.data(), nfft * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
      // H = DFT(h) / nfft  (the 1/nfft of the forward/inverse round trip is folded in here)
      launch_fft_pow2(f->d_H, f->d_H, nfft, 1, f->d_tw, false, true, 1.0f / (float)nfft, c->stream);
      CKL(1);
      static const char* os_v1 = getenv("AE_FIR_OS_V1");   // developer switch: K4 (two warps per segment) for 1024-point blocks too
      if (nfft == 1024 && !os_v1) {
        std::vector<float2> tw, hi, lo;
        chain_x2_tables(1024, h.data(), 1, tw, hi, lo);
        TRY(dev_alloc(c, tw.size() * sizeof(float2), &p)); f->d_x2tw = (float2*)p;
        CK(cudaMemcpyAsync(f->d_x2tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));   // tw is a local of this block
      }
    }
    CK(cudaStreamSynchronize(c->stream));   // also keeps the host staging vectors alive until the copies are done
    return AE_OK;
  };
  const ae_status st = build();
  if (st != AE_OK) { ae_fir_destroy(f); return st; }
  *out = f;
  return AE_OK;
}
ae_status ae_fir_destroy(ae_fir* f) {
  if (!f) return AE_OK;
  cudaSetDevice(f->c->dev);
  dev_free(f->c, f->d_taps); dev_free(f->c, f->d_hist); dev_free(f->c, f->d_hist2); dev_free(f->c, f->d_H); dev_free(f->c, f->d_x2tw);
  delete f;
  return AE_OK;
}
size_t ae_fir_ntaps(const ae_fir* f) { return f ? f->ntaps : 0; }
size_t ae_fir_block_hop(const ae_fir* f) {
  if (!f) return 0;
  return f->mode == AE_FIR_OVERLAP_SAVE ? f->nfft - f->ntaps + 1 : 1;
}
ae_status ae_fir_reset(ae_fir* f) {
  if (!f) return fail(AE_EARG, "null");
  cudaSetDevice(f->c->dev);
  CK(cudaMemsetAsync(f->d_hist, 0, f->tp * sizeof(float2), f->c->stream));
  return AE_OK;
}

ae_status ae_fir_exec(ae_fir* f, ae_vec* in, ae_vec* out, size_t frame_len) {
  if (!f || !in || !out) return fail(AE_EARG, "null");
  if (in->len != out->len) return fail(AE_ELEN, "Vectors must have same length");
  Ctx* c = f->c;
  cudaSetDevice(c->dev);
  const size_t n = in->len;
  if (n == 0) return AE_OK;
  TRY(before_read(in));
  TRY(before_write(out));
  const float2* x = vptr(in);
  float2* y = vptr(out);
  void* tmp = nullptr;
  if (in->a == out->a) {  // blocks read a halo other blocks overwrite: filter from a private copy
    TRY(dev_alloc(c, n * sizeof(float2), &tmp));
    CK(cudaMemcpyAsync(tmp, x, n * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
    x = (const float2*)tmp;
  }
  int mode = f->mode;
  if (mode == AE_FIR_DIRECT && frame_len && (frame_len % 8) != 0) {
    if (!f->d_H) { dev_free(c, tmp); return fail(AE_EARG, "direct-form FIR needs frame_len % 8 == 0"); }
    mode = AE_FIR_OVERLAP_SAVE;
  }
  const float2* hist = frame_len ? nullptr : f->d_hist;
  if (mode == AE_FIR_DIRECT) launch_fir_direct(x, y, n, f->d_taps, f->tp, hist, frame_len, c->sm_count, c->stream, f->h_taps.data());
  else if (f->d_x2tw) launch_fir_os_x2(x, y, n, f->d_H, f->d_x2tw, f->ntaps, hist, frame_len, c->stream);
  else launch_fir_overlap_save(x, y, n, f->d_H, f->d_tw, f->nfft, f->ntaps, hist, frame_len, c->stream);
  CKL(1);
  if (!frame_len && f->tp > 1) {  // carry the last tp-1 inputs
    const size_t hl = (size_t)f->tp - 1;
    if (n >= hl) {
      CK(cudaMemcpyAsync(f->d_hist, x + (n - hl), hl * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
    } else {
      CK(cudaMemcpyAsync(f->d_hist2, f->d_hist + n, (hl - n) * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaMemcpyAsync(f->d_hist2 + (hl - n), x, n * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
      std::swap(f->d_hist, f->d_hist2);
    }
  }
  dev_free(c, tmp);
  return AE_OK;
}

// =================================================================================================
// sampling
// =================================================================================================
ae_status ae_interpolate(ae_vec* src, ae_vec* dst, size_t n_between, int compat) {
  if (!src || !dst) return fail(AE_EARG, "null");
  if (src->len == 0) return fail(AE_EARG, "called `Option::unwrap()` on a `None` value");  // src/sampling.rs:23
  if (src->a == dst->a) return fail(AE_EARG, "interpolate: dst aliases src");
  if (n_between >= (1u << 30)) return fail(AE_EARG, "n_between too large");
  cudaSetDevice(src->c->dev);
  const size_t n_out = (src->len - 1) * (n_between + 1) + 1;
  TRY(before_read(src));
  TRY(before_write(dst));
  TRY(vec_reserve(dst, dst->len + n_out));
  launch_interpolate(vptr(src), src->len, vptr(dst) + dst->len, n_between, compat, src->c->stream);
  CKL(1);
  dst->len += n_out;  // appended, like Vec::push
  return AE_OK;
}
static ae_status downsample_impl(ae_vec* src, ae_vec* dst, int strict) {
  if (!src || !dst) return fail(AE_EARG, "null");
  if (dst->len == 0) return fail(AE_EARG, "attempt to divide by zero");                       // src/sampling.rs:38
  if (strict && (src->len % dst->len) != 0) return fail(AE_ELEN, "Only even decimations are supported");  // :32
  if (src->a == dst->a) return fail(AE_EARG, "downsample: dst aliases src");
  const size_t dec = src->len / dst->len;
  if (dst->len > 0 && (dst->len - 1) * dec >= src->len && src->len > 0 && dec > 0) return fail(AE_EIDX, "index out of bounds");
  if (src->len == 0) return fail(AE_EIDX, "index out of bounds");
  cudaSetDevice(src->c->dev);
  TRY(before_read(src));
  TRY(before_write(dst));
  launch_downsample_cf32(vptr(src), vptr(dst), dst->len, dec, src->c->stream);
  CKL(1);
  return AE_OK;
}
ae_status ae_downsample(ae_vec* src, ae_vec* dst, int strict) { return downsample_impl(src, dst, strict); }
ae_status ae_downsample_sb(ae_vec* src, ae_vec* dst, int strict) {
  // step_by variant (src/sampling.rs:49-62): zip stops at the shorter side, step_by(0) panics
  if (src && dst && dst->len && src->len / dst->len == 0) return fail(AE_EARG, "assertion failed: step != 0");
  return downsample_impl(src, dst, strict);
}
ae_status ae_downsample_bits(ae_bits* src, ae_bits* dst, int strict) {
  if (!src || !dst) return fail(AE_EARG, "null");
  if (dst->len == 0) return fail(AE_EARG, "attempt to divide by zero");
  if (strict && (src->len % dst->len) != 0) return fail(AE_ELEN, "Only even decimations are supported");
  if (src->len == 0) return fail(AE_EIDX, "index out of bounds");
  const size_t dec = src->len / dst->len;
  cudaSetDevice(src->c->dev);
  launch_downsample_u8(bptr(src), bptr(dst), dst->len, dec, src->c->stream);
  CKL(1);
  return AE_OK;
}

}  // extern "C"

// =================================================================================================
// modulation
// =================================================================================================
struct ae_mod {
  Ctx* c;
  ModTable tab;
};

extern "C" {

ae_status ae_mod_create(const ae_cf32* table, size_t table_len, ae_mod** out) {
  if (!out || !table) return fail(AE_EARG, "null");
  if (table_len != 2 && table_len != 4) return fail(AE_EARG, "Modulation is implemented for [cf32; 2] and [cf32; 4] only");
  Ctx* c;
  TRY(get_ctx(&c));
  ae_mod* m = new ae_mod;
  m->c = c;
  m->tab.len = (int)table_len;
  for (size_t i = 0; i < 4; ++i) m->tab.t[i] = i < table_len ? make_float2(table[i].re, table[i].im) : make_float2(0.f, 0.f);
  m->tab.generic_qpsk = table_len == 4 && table[0].re == 1.f && table[0].im == 1.f && table[1].re == -1.f && table[1].im == 1.f &&
                        table[2].re == 1.f && table[2].im == -1.f && table[3].re == -1.f && table[3].im == -1.f;
  *out = m;
  return AE_OK;
}
ae_status ae_mod_bpsk(ae_mod** out) {
  const ae_cf32 t[2] = {{1.0f, 1.0f}, {-1.0f, -1.0f}};  // GENERIC_BPSK_TABLE src/modulation.rs:77
  return ae_mod_create(t, 2, out);
}
ae_status ae_mod_qpsk(ae_mod** out) {
  const ae_cf32 t[4] = {{1.0f, 1.0f}, {-1.0f, 1.0f}, {1.0f, -1.0f}, {-1.0f, -1.0f}};  // src/modulation.rs:87-92
  return ae_mod_create(t, 4, out);
}
ae_status ae_mod_destroy(ae_mod* m) { delete m; return AE_OK; }
size_t ae_mod_bits_per_symbol(const ae_mod* m) { return m ? (m->tab.len == 2 ? 1 : 2) : 0; }

static ae_status modulate_impl(ae_mod* m, ae_bits* bits, ae_vec* out, bool collect) {
  if (!m || !bits || !out) return fail(AE_EARG, "null");
  const size_t bps = ae_mod_bits_per_symbol(m);
  cudaSetDevice(m->c->dev);
  const size_t nsym_full = bits->len / bps;
  const bool ragged = (bits->len % bps) != 0;
  size_t n_out;
  if (collect) {
    // chunks(BPS) yields a short last chunk; QPSK index() then reads bits[1] out of bounds (:24)
    if (ragged) return fail(AE_EIDX, "index out of bounds: the len is 1 but the index is 1");
    n_out = nsym_full;
    TRY(before_write(out));
    TRY(flush_vec(out));
    TRY(vec_reserve(out, n_out));
    out->len = n_out;
  } else {
    TRY(before_write(out));
    TRY(flush_vec(out));
    n_out = std::min(out->len, nsym_full + (ragged ? 1 : 0));  // zip truncates (:123-131)
    if (ragged && n_out > nsym_full) return fail(AE_EIDX, "index out of bounds: the len is 1 but the index is 1");
  }
  launch_modulate(m->tab, bptr(bits), bits->len, vptr(out), n_out, m->c->d_err, m->c->stream);
  if (n_out) CKL(1);
  return AE_OK;
}
ae_status ae_mod_modulate(ae_mod* m, ae_bits* bits, ae_vec* out) { return modulate_impl(m, bits, out, true); }
ae_status ae_mod_modulate_into(ae_mod* m, ae_bits* bits, ae_vec* out) { return modulate_impl(m, bits, out, false); }

ae_status ae_mod_demod(ae_mod* m, ae_vec* symbols, ae_bits* out, int compat) {
  if (!m || !symbols || !out) return fail(AE_EARG, "null");
  const size_t bps = ae_mod_bits_per_symbol(m);
  cudaSetDevice(m->c->dev);
  TRY(before_read(symbols));
  const size_t add = symbols->len * bps;
  TRY(bits_reserve(out, out->len + add));
  launch_demod(m->tab, vptr(symbols), symbols->len, bptr(out) + out->len, compat, m->c->stream);
  if (symbols->len) CKL(1);
  out->len += add;  // output.push / extend (:51-53, :142)
  return AE_OK;
}

}  // extern "C"

// =================================================================================================
// noise
// =================================================================================================
struct ae_awgn {
  Ctx* c;
  float power, scale;
  uint64_t seed, stream_id, offset;
};

extern "C" {

ae_status ae_awgn_create(float power, uint64_t seed, ae_awgn** out) {
  if (!out) return fail(AE_EARG, "null");
  Ctx* c;
  TRY(get_ctx(&c));
  ae_awgn* g = new ae_awgn;
  g->c = c; g->power = power; g->scale = sqrtf(power);  // src/noise.rs:35
  g->seed = seed; g->stream_id = 0; g->offset = 0;
  *out = g;
  return AE_OK;
}
ae_status ae_awgn_generator(ae_awgn** out) { return ae_awgn_create(1.0f, 815ull, out); }  // src/noise.rs:6-11
ae_status ae_awgn_destroy(ae_awgn* g) { delete g; return AE_OK; }
ae_status ae_awgn_set_power(ae_awgn* g, float power) {
  if (!g) return fail(AE_EARG, "null");
  g->power = power; g->scale = sqrtf(power);
  return AE_OK;
}
ae_status ae_awgn_set_stream_id(ae_awgn* g, uint64_t id) {
  if (!g) return fail(AE_EARG, "null");
  g->stream_id = id;
  return AE_OK;
}
ae_status ae_awgn_seek(ae_awgn* g, uint64_t off) {
  if (!g) return fail(AE_EARG, "null");
  g->offset = off;
  return AE_OK;
}
uint64_t ae_awgn_tell(const ae_awgn* g) { return g ? g->offset : 0; }

ae_status ae_awgn_fill(ae_awgn* g, ae_vec* target) {
  if (!g || !target) return fail(AE_EARG, "null");
  cudaSetDevice(g->c->dev);
  TRY(before_write(target));
  TRY(flush_vec(target));
  const size_t n = target->cap - target->len;  // while len < capacity { push(next()) }
  launch_awgn_fill(vptr(target) + target->len, n, g->scale, g->seed, g->stream_id, g->offset, g->c->stream);
  if (n) CKL(1);
  target->len = target->cap;
  g->offset += n;
  return AE_OK;
}
ae_status ae_awgn_apply(ae_awgn* g, ae_vec* signal, int compat) {
  if (!g || !signal) return fail(AE_EARG, "null");
  cudaSetDevice(g->c->dev);
  TRY(before_write(signal));
  TRY(flush_vec(signal));
  launch_awgn_apply(vptr(signal), signal->len, g->scale, compat == AE_COMPAT_REFERENCE ? 1 : 0, g->seed, g->stream_id,
                    g->offset, g->c->stream);
  if (signal->len) CKL(1);
  g->offset += signal->len;
  return AE_OK;
}
ae_status ae_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  if (!ctr || !key || !out) return fail(AE_EARG, "null");
  uint32_t o[4];
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], o);
  for (int i = 0; i < 4; ++i) out[i] = o[i];
  return AE_OK;
}

// =================================================================================================
// sequence
// =================================================================================================
ae_status ae_mseq_expand(uint64_t seed, size_t len, ae_bits* out) {
  if (!out) return fail(AE_EARG, "null");
  if (len > 64) return fail(AE_EARG, "attempt to shift right with overflow");  // src/sequence.rs:20, i >= 64
  cudaSetDevice(out->c->dev);
  TRY(bits_reserve(out, len));
  out->len = len;
  launch_expand(seed, len, bptr(out), out->c->stream);
  if (len) CKL(1);
  return AE_OK;
}
ae_status ae_mseq_generate(const uint8_t* init_host, size_t n_init, const uint32_t* back, size_t n_back, size_t len,
                           ae_bits* out) {
  if (!out || (!init_host && n_init) || (!back && n_back)) return fail(AE_EARG, "null");
  Ctx* c = out->c;
  cudaSetDevice(c->dev);
  const size_t total = std::max(len, n_init);  // generate() never truncates init (src/sequence.rs:48)
  TRY(bits_reserve(out, total));
  out->len = total;
  if (len > n_init) {
    uint32_t deg = 0;
    for (size_t t = 0; t < n_back; ++t) {
      if (back[t] == 0) return fail(AE_EIDX, "index out of bounds: generator reads the element it is producing");
      deg = std::max(deg, back[t]);
    }
    if (deg > n_init) return fail(AE_EIDX, "attempt to subtract with overflow: init shorter than the deepest tap");
    if (deg > 64) return fail(AE_EARG, "tap offsets above 64 are not supported");
    if (deg == 0) {
      // no taps: every generated element is (empty sum) % 2 = 0
      CK(cudaMemsetAsync(bptr(out) + n_init, 0, len - n_init, c->stream));
    } else {
      const size_t base = n_init - deg;
      uint64_t poly_low = 0, state = 0;
      for (size_t t = 0; t < n_back; ++t) poly_low ^= 1ull << (deg - back[t]);
      for (uint32_t j = 0; j < deg; ++j) state |= (uint64_t)(init_host[base + j] & 1u) << j;
      launch_mseq(state, poly_low, (int)deg, len - base, bptr(out) + base, c->stream);
      CKL(1);
    }
  }
  if (n_init) {  // init is returned verbatim, whatever byte values it holds
    CK(cudaMemcpyAsync(bptr(out), init_host, n_init, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return AE_OK;
}

// =================================================================================================
// statistics
// =================================================================================================
ae_status ae_stats_alloc(ae_stats** d) {
  if (!d) return fail(AE_EARG, "null");
  Ctx* c;
  TRY(get_ctx(&c));
  CK(cudaMalloc((void**)d, sizeof(ae_stats)));
  CK(cudaMemsetAsync(*d, 0, sizeof(ae_stats), c->stream));
  return AE_OK;
}
ae_status ae_stats_free(ae_stats* d) {
  if (d) CK(cudaFree(d));
  return AE_OK;
}
ae_status ae_stats_zero(ae_stats* d) {
  Ctx* c;
  TRY(get_ctx(&c));
  if (!d) return fail(AE_EARG, "null");
  CK(cudaMemsetAsync(d, 0, sizeof(ae_stats), c->stream));
  return AE_OK;
}
ae_status ae_stats_read(const ae_stats* d, ae_stats* host) {
  Ctx* c;
  TRY(get_ctx(&c));
  if (!d || !host) return fail(AE_EARG, "null");
  CK(cudaMemcpyAsync(host, d, sizeof(ae_stats), cudaMemcpyDeviceToHost, c->stream));
  return ae_sync();
}
}  // extern "C"

// -------------------------------------------------------------------------------------------------
// The one collective of the path: ae_stats summed over the GPUs with NCCL.  NCCL is bound at
// run time (dlopen of the image's libnccl.so.2, or the copy torch already loaded), so the library has no
// link-time dependency on it; the handful of declarations below are NCCL's stable C ABI.
// -------------------------------------------------------------------------------------------------
namespace {
struct NcclId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommInitAll)(NcclComm*, int, const int*) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
constexpr int kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values
NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // torch's copy when it is already in the process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    bool all = true;
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) { all = false; api.why = std::string("libnccl lacks ") + n; } return p; };
    api.GetUniqueId = (int (*)(NcclId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclId, int))sym("ncclCommInitRank");
    api.CommInitAll = (int (*)(NcclComm*, int, const int*))sym("ncclCommInitAll");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclAllReduce");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    api.ok = all;
  });
  return api;
}
ae_status nccl_fail(const NcclApi& n, int rc, const char* what) {
  return fail(AE_ENCCL, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}
#define NCK(call, what)                                  \
  do {                                                   \
    int r__ = (call);                                    \
    if (r__ != 0) return nccl_fail(n, r__, what);        \
  } while (0)
}  // namespace

struct ae_comm {
  Ctx* c;
  NcclComm comm;
  int nranks, rank;
};

namespace {
// both reductions of one ae_stats (two u64 counters, two f64 sums), to be called inside an NCCL group
ae_status stats_allreduce_enqueue(NcclApi& n, ae_stats* d, ae_comm* cm) {
  cudaSetDevice(cm->c->dev);
  NCK(n.AllReduce(&d->bit_errors, &d->bit_errors, 2, kNcclUint64, kNcclSum, cm->comm, cm->c->stream), "ncclAllReduce(u64)");
  NCK(n.AllReduce(&d->err_pow, &d->err_pow, 2, kNcclFloat64, kNcclSum, cm->comm, cm->c->stream), "ncclAllReduce(f64)");
  return AE_OK;
}
}  // namespace

extern "C" {

ae_status ae_comm_unique_id(uint8_t id_out[AE_COMM_ID_BYTES]) {
  if (!id_out) return fail(AE_EARG, "null");
  NcclApi& n = nccl();
  if (!n.ok) return fail(AE_ENCCL, n.why);
  NcclId id;
  NCK(n.GetUniqueId(&id), "ncclGetUniqueId");
  static_assert(sizeof(NcclId) == AE_COMM_ID_BYTES, "NCCL unique id is 128 bytes");
  std::memcpy(id_out, &id, sizeof(id));
  return AE_OK;
}
ae_status ae_comm_init_rank(const uint8_t id_in[AE_COMM_ID_BYTES], int nranks, int rank, ae_comm** out) {
  if (!id_in || !out) return fail(AE_EARG, "null");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(AE_EARG, "bad rank / nranks");
  NcclApi& n = nccl();
  if (!n.ok) return fail(AE_ENCCL, n.why);
  Ctx* c;
  TRY(get_ctx(&c));
  NcclId id;
  std::memcpy(&id, id_in, sizeof(id));
  NcclComm comm = nullptr;
  NCK(n.CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
  *out = new ae_comm{c, comm, nranks, rank};
  return AE_OK;
}
ae_status ae_comm_init_all(int ndev, ae_comm** comms_out) {
  if (!comms_out || ndev < 1) return fail(AE_EARG, "null / ndev < 1");
  NcclApi& n = nccl();
  if (!n.ok) return fail(AE_ENCCL, n.why);
  const int keep = t_dev;
  std::vector<Ctx*> ctx(ndev);
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; ++i) {
    ae_status st = ae_init(i);
    if (st == AE_OK) st = get_ctx(&ctx[i]);
    if (st != AE_OK) { if (keep >= 0) ae_init(keep); return st; }
    devs[i] = i;
  }
  if (keep >= 0) ae_init(keep);
  std::vector<NcclComm> comms(ndev, nullptr);
  NCK(n.CommInitAll(comms.data(), ndev, devs.data()), "ncclCommInitAll");
  for (int i = 0; i < ndev; ++i) comms_out[i] = new ae_comm{ctx[i], comms[i], ndev, i};
  return AE_OK;
}
ae_status ae_comm_destroy(ae_comm* cm) {
  if (!cm) return AE_OK;
  NcclApi& n = nccl();
  cudaSetDevice(cm->c->dev);
  cudaStreamSynchronize(cm->c->stream);
  int rc = n.ok ? n.CommDestroy(cm->comm) : 0;
  delete cm;
  if (rc != 0) return nccl_fail(n, rc, "ncclCommDestroy");
  return AE_OK;
}
ae_status ae_comm_info(const ae_comm* cm, int* nranks, int* rank, int* device) {
  if (!cm) return fail(AE_EARG, "null");
  if (nranks) *nranks = cm->nranks;
  if (rank) *rank = cm->rank;
  if (device) *device = cm->c->dev;
  return AE_OK;
}
ae_status ae_stats_allreduce(ae_stats* d, ae_comm* cm) {
  if (!d || !cm) return fail(AE_EARG, "null");
  NcclApi& n = nccl();
  if (!n.ok) return fail(AE_ENCCL, n.why);
  NCK(n.GroupStart(), "ncclGroupStart");
  ae_status st = stats_allreduce_enqueue(n, d, cm);
  const int rc = n.GroupEnd();
  if (st != AE_OK) return st;
  if (rc != 0) return nccl_fail(n, rc, "ncclGroupEnd");
  return AE_OK;
}
ae_status ae_stats_allreduce_all(ae_stats** d, ae_comm** cm, int ndev) {
  if (!d || !cm || ndev < 1) return fail(AE_EARG, "null / ndev < 1");
  for (int i = 0; i < ndev; ++i)
    if (!d[i] || !cm[i]) return fail(AE_EARG, "null");
  NcclApi& n = nccl();
  if (!n.ok) return fail(AE_ENCCL, n.why);
  const int keep = t_dev;
  NCK(n.GroupStart(), "ncclGroupStart");
  ae_status st = AE_OK;
  for (int i = 0; i < ndev && st == AE_OK; ++i) st = stats_allreduce_enqueue(n, d[i], cm[i]);
  const int rc = n.GroupEnd();
  if (keep >= 0) cudaSetDevice(g_ctx[keep]->dev);
  if (st != AE_OK) return st;
  if (rc != 0) return nccl_fail(n, rc, "ncclGroupEnd");
  return AE_OK;
}

ae_status ae_count_bit_errors(ae_bits* a, ae_bits* b, ae_stats* d) {
  if (!a || !b || !d) return fail(AE_EARG, "null");
  if (a->len != b->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(a->c->dev);
  launch_bit_errors(bptr(a), bptr(b), a->len, d, a->c->sm_count, a->c->stream);
  if (a->len) CKL(1);
  return AE_OK;
}
ae_status ae_evm_accumulate(ae_vec* act, ae_vec* ref, ae_stats* d) {
  if (!act || !ref || !d) return fail(AE_EARG, "null");
  if (act->len != ref->len) return fail(AE_ELEN, "Input slices/vectors must be same length");  // src/lib.rs:34
  cudaSetDevice(act->c->dev);
  TRY(before_read(act));
  TRY(before_read(ref));
  launch_evm_acc(vptr(act), vptr(ref), act->len, d, act->c->sm_count, act->c->stream);
  if (act->len) CKL(1);
  return AE_OK;
}

static ae_status vecstats_impl(Ctx* c, const void* p, size_t n, bool cplx, ae_vecstats* host_out) {
  if (!host_out) return fail(AE_EARG, "null");
  if (n == 0) return fail(AE_ELEN, "VecStats of an empty vector");
  cudaSetDevice(c->dev);
  void* scratch = nullptr;
  const size_t bytes = vecstats_scratch_bytes(c->sm_count);
  TRY(dev_alloc(c, bytes, &scratch));
  const int launches = launch_vecstats(p, n, cplx, scratch, c->sm_count, c->stream);
  ae_status st = AE_OK;
  if (cudaGetLastError() != cudaSuccess) st = fail(AE_ECUDA, "VecStats launch failed");
  else {
    g_launches += launches;
    const char* res = (const char*)scratch + bytes - sizeof(ae_vecstats);
    if (cudaMemcpyAsync(host_out, res, sizeof(ae_vecstats), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      st = fail(AE_ECUDA, "VecStats read-back failed");
    else st = sync_ctx(c);
  }
  dev_free(c, scratch);
  return st;
}
ae_status ae_vec_stats(ae_vec* v, ae_vecstats* host_out) {
  if (!v) return fail(AE_EARG, "null");
  TRY(before_read(v));
  return vecstats_impl(v->c, vptr(v), v->len, true, host_out);
}

// =================================================================================================
// fused chains
// =================================================================================================
ae_status ae_modem_fused(ae_mod* m, ae_awgn* g, ae_bits* bits_in, ae_bits* bits_out, ae_stats* d, int compat) {
  if (!m || !g || !bits_in || !bits_out) return fail(AE_EARG, "null");
  const size_t bps = ae_mod_bits_per_symbol(m);
  if (bits_in->len % bps) return fail(AE_EIDX, "index out of bounds: the len is 1 but the index is 1");
  cudaSetDevice(m->c->dev);
  TRY(bits_reserve(bits_out, bits_in->len));
  bits_out->len = bits_in->len;
  launch_modem_fused(m->tab, bptr(bits_in), bits_in->len, bptr(bits_out), g->scale, compat == AE_COMPAT_REFERENCE ? 1 : 0,
                     g->seed, g->stream_id, g->offset, compat, d, m->c->d_err, m->c->sm_count, m->c->stream);
  if (bits_in->len) CKL(1);
  g->offset += bits_in->len / bps;
  return AE_OK;
}

}  // extern "C"

struct ae_chain {
  Ctx* c;
  size_t n, ntaps;
  int scale_kind, compat;
  float x, s;
  bool fused;
  float2 *d_tw, *d_window, *d_taps;
  bool x2;                                // K14b applies (N = 1024, <= 64 taps)
  float2 *d_x2tw, *d_taps_hi, *d_taps_lo;
  ae_fft* fft;
  ae_fir* fir;
  ae_mod* qpsk;
  // host pipeline
  float2* d_in[3];
  uint8_t* d_out[3];
  size_t chunk_frames;
};

extern "C" {

ae_status ae_chain_create(size_t fft_len, const ae_cf32* taps_host, size_t ntaps, int scale_kind, float x, int compat,
                          ae_chain** out) {
  if (!out || !taps_host) return fail(AE_EARG, "null");
  if (fft_len == 0 || ntaps == 0) return fail(AE_EARG, "empty chain");
  if (scale_kind < 0 || scale_kind > 3) return fail(AE_EARG, "bad scale kind");
  Ctx* c;
  TRY(get_ctx(&c));
  ae_chain* ch = new ae_chain;
  std::memset(ch, 0, sizeof(*ch));
  ch->c = c; ch->n = fft_len; ch->ntaps = ntaps; ch->scale_kind = scale_kind; ch->x = x; ch->compat = compat;
  ch->s = scale_factor(scale_kind, fft_len, x);
  ch->fused = chain_fused_supported(fft_len, ntaps);
  ae_status st = ae_fft_create(fft_len, &ch->fft);
  if (st == AE_OK) { ae_fft_set_compat(ch->fft, compat); st = ae_fir_create(taps_host, ntaps, AE_FIR_AUTO, &ch->fir); }
  if (st == AE_OK) st = ae_mod_qpsk(&ch->qpsk);
  if (st != AE_OK) { ae_chain_destroy(ch); return st; }
  if (ch->fused) {
    // every failure below goes through ae_chain_destroy (nothing leaks)
    auto fail_out = [&](ae_status e) { ae_chain_destroy(ch); return e; };
    st = get_thread_twiddles(c, fft_len, &ch->d_tw);
    if (st != AE_OK) return fail_out(st);
    // window[m] = s * sum_k h[k] exp(-sgn 2 pi i m k / N), sgn = exponent sign of Cfft::fwd
    const double sgn = (compat == AE_COMPAT_REFERENCE) ? +1.0 : -1.0;
    std::vector<float2> w(fft_len), h(ntaps);
    for (size_t m = 0; m < fft_len; ++m) {
      double ar = 0, ai = 0;
      for (size_t k = 0; k < ntaps; ++k) {
        const double a = -sgn * 2.0 * M_PI * (double)((m * k) % fft_len) / (double)fft_len;
        const double cr = std::cos(a), ci = std::sin(a);
        ar += taps_host[k].re * cr - taps_host[k].im * ci;
        ai += taps_host[k].re * ci + taps_host[k].im * cr;
      }
      w[m] = make_float2((float)(ar * (double)ch->s), (float)(ai * (double)ch->s));
    }
    for (size_t k = 0; k < ntaps; ++k) h[k] = make_float2(taps_host[k].re, taps_host[k].im);
    auto upload = [&](const std::vector<float2>& v, float2** dst) -> ae_status {
      void* p = nullptr;
      TRY(dev_alloc(c, v.size() * sizeof(float2), &p));
      *dst = (float2*)p;
      CK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(float2), cudaMemcpyHostToDevice, c->stream));
      return AE_OK;
    };
    st = upload(w, &ch->d_window);
    if (st == AE_OK) st = upload(h, &ch->d_taps);
    static const char* force_v1 = getenv("AE_CHAIN_V1");   // developer switch: K14 (chain.cu) instead of K14b
    ch->x2 = chain_x2_supported(fft_len, ntaps) && !force_v1;
    if (st == AE_OK && ch->x2) {
      std::vector<float2> tw, hi, lo;
      chain_x2_tables(fft_len, h.data(), ntaps, tw, hi, lo);
      st = upload(tw, &ch->d_x2tw);
      if (st == AE_OK) st = upload(hi, &ch->d_taps_hi);
      if (st == AE_OK) st = upload(lo, &ch->d_taps_lo);
    }
    if (st == AE_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) st = fail(AE_ECUDA, "chain table upload failed");
    if (st != AE_OK) return fail_out(st);
  }
  *out = ch;
  return AE_OK;
}
ae_status ae_chain_destroy(ae_chain* ch) {
  if (!ch) return AE_OK;
  cudaSetDevice(ch->c->dev);
  ae_fft_destroy(ch->fft); ae_fir_destroy(ch->fir); ae_mod_destroy(ch->qpsk);
  dev_free(ch->c, ch->d_window); dev_free(ch->c, ch->d_taps);
  dev_free(ch->c, ch->d_x2tw); dev_free(ch->c, ch->d_taps_hi); dev_free(ch->c, ch->d_taps_lo);
  for (int i = 0; i < 3; ++i) { if (ch->d_in[i]) cudaFree(ch->d_in[i]); if (ch->d_out[i]) cudaFree(ch->d_out[i]); }
  delete ch;
  return AE_OK;
}

}  // extern "C"
// one launch of the fused chain: K14b when it applies and the bit buffer is 4-byte aligned, K14 otherwise
static void chain_launch(ae_chain* ch, const float2* x, uint8_t* bits, size_t frames, cudaStream_t st) {
  const bool inverse = (ch->compat == AE_COMPAT_REFERENCE);
  if (ch->x2 && ((uintptr_t)bits % 4) == 0)
    launch_chain_x2(x, bits, frames, ch->d_window, ch->d_x2tw, ch->d_taps_hi, ch->d_taps_lo, ch->ntaps, inverse, ch->s, ch->compat, st);
  else
    launch_chain_fused(x, bits, ch->n, frames, ch->d_window, ch->d_taps, ch->ntaps, ch->d_tw, inverse, ch->s, ch->compat, st);
}
extern "C" {

ae_status ae_chain_exec_unfused(ae_chain* ch, ae_vec* in, ae_bits* bits_out, ae_vec* symbols_out) {
  if (!ch || !in || !bits_out) return fail(AE_EARG, "null");
  if (in->len % ch->n) return fail(AE_ELEN, "Input and FFT must be the same length");
  cudaSetDevice(ch->c->dev);
  const size_t frames = in->len / ch->n;
  ae_vec *X = nullptr, *Y = symbols_out;
  TRY(ae_vec_alloc(in->len, in->len, &X));
  ae_status st = AE_OK;
  if (!Y) st = ae_vec_alloc(in->len, in->len, &Y);
  else if (Y->len != in->len) st = fail(AE_ELEN, "Vectors must have same length");
  if (st == AE_OK) st = ae_fft_exec(ch->fft, AE_FFT_FWD, in, X, ch->scale_kind, ch->x, frames);
  if (st == AE_OK) st = ae_fir_exec(ch->fir, X, Y, ch->n);
  if (st == AE_OK) { bits_out->len = 0; st = ae_mod_demod(ch->qpsk, Y, bits_out, ch->compat); }
  ae_vec_free(X);
  if (Y != symbols_out) ae_vec_free(Y);
  return st;
}

ae_status ae_chain_exec(ae_chain* ch, ae_vec* in, ae_bits* bits_out) {
  if (!ch || !in || !bits_out) return fail(AE_EARG, "null");
  if (in->len % ch->n) return fail(AE_ELEN, "Input and FFT must be the same length");
  if (!ch->fused) return ae_chain_exec_unfused(ch, in, bits_out, nullptr);
  // the fused kernels write the two bytes of a symbol with one 16-bit (K14) or 32-bit (K14b) store: a wrapped bit
  // buffer at an odd address takes the composition of the stand-alone kernels
  if ((reinterpret_cast<uintptr_t>(bits_out->a->p) + bits_out->off) & 1) return ae_chain_exec_unfused(ch, in, bits_out, nullptr);
  Ctx* c = ch->c;
  cudaSetDevice(c->dev);
  const size_t frames = in->len / ch->n;
  TRY(before_read(in));
  TRY(bits_reserve(bits_out, 2 * in->len));
  bits_out->len = 2 * in->len;
  chain_launch(ch, vptr(in), bptr(bits_out), frames, c->stream);
  if (frames) CKL(1);
  return AE_OK;
}

ae_status ae_chain_exec_host(ae_chain* ch, const ae_cf32* host_in, size_t n_samples, uint8_t* host_bits) {
  if (!ch || (!host_in && n_samples) || (!host_bits && n_samples)) return fail(AE_EARG, "null");
  if (n_samples % ch->n) return fail(AE_ELEN, "Input and FFT must be the same length");
  if (!ch->fused) return fail(AE_EARG, "host pipeline needs the fused chain (power-of-two FFT length 256..4096)");
  Ctx* c = ch->c;
  cudaSetDevice(c->dev);
  const size_t frames = n_samples / ch->n;
  // 3-deep pipeline of (H2D, kernel, D2H) on three streams; chunk = 64 MiB of input
  if (!ch->chunk_frames) {
    ch->chunk_frames = std::max<size_t>(1, ((size_t)64 << 20) / (ch->n * sizeof(float2)));
    for (int i = 0; i < 3; ++i) {
      CK(cudaMalloc((void**)&ch->d_in[i], ch->chunk_frames * ch->n * sizeof(float2)));
      CK(cudaMalloc((void**)&ch->d_out[i], ch->chunk_frames * ch->n * 2));
    }
  }
  // order the pipeline after work already queued on the context stream
  cudaEvent_t ev;
  CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  auto run = [&]() -> ae_status {
    CK(cudaEventRecord(ev, c->stream));
    for (int i = 0; i < 3; ++i) CK(cudaStreamWaitEvent(c->pipe[i], ev, 0));
    size_t done = 0;
    int slot = 0;
    while (done < frames) {
      const size_t fc = std::min(ch->chunk_frames, frames - done);
      cudaStream_t st = c->pipe[slot];
      CK(cudaMemcpyAsync(ch->d_in[slot], host_in + done * ch->n, fc * ch->n * sizeof(float2), cudaMemcpyHostToDevice, st));
      chain_launch(ch, ch->d_in[slot], ch->d_out[slot], fc, st);
      CKL(1);
      CK(cudaMemcpyAsync(host_bits + 2 * done * ch->n, ch->d_out[slot], fc * ch->n * 2, cudaMemcpyDeviceToHost, st));
      done += fc;
      slot = (slot + 1) % 3;
    }
    for (int i = 0; i < 3; ++i) {
      CK(cudaEventRecord(ev, c->pipe[i]));
      CK(cudaStreamWaitEvent(c->stream, ev, 0));
    }
    for (int i = 0; i < 3; ++i) CK(cudaStreamSynchronize(c->pipe[i]));
    return AE_OK;
  };
  const ae_status st = run();
  cudaEventDestroy(ev);      // on every path
  return st;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// CUDA graphs over the context stream (launch-bound shapes)
// -------------------------------------------------------------------------------------------------
struct ae_graph {
  Ctx* c;
  cudaGraph_t graph;
  cudaGraphExec_t exec;
};
extern "C" {
ae_status ae_graph_begin(void) {
  Ctx* c;
  TRY(get_ctx(&c));
  // VecOps recorded before the bracket run now, once - not as part of every replay
  {
    std::vector<ae_vec*> snap = c->pending;
    for (ae_vec* v : snap) TRY(flush_vec(v));
  }
  CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
  return AE_OK;
}
ae_status ae_graph_end(ae_graph** out) {
  if (!out) return fail(AE_EARG, "null");
  Ctx* c;
  TRY(get_ctx(&c));
  // VecOps recorded inside the bracket and not yet run belong to the graph
  ae_status fst = AE_OK;
  {
    std::vector<ae_vec*> snap = c->pending;
    for (ae_vec* v : snap)
      if (fst == AE_OK) fst = flush_vec(v);
  }
  cudaGraph_t g = nullptr;
  CK(cudaStreamEndCapture(c->stream, &g));       // always leave capture mode
  if (fst != AE_OK) { if (g) cudaGraphDestroy(g); return fst; }
  cudaGraphExec_t ex = nullptr;
  const cudaError_t e = cudaGraphInstantiate(&ex, g, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(AE_ECUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e)); }
  *out = new ae_graph{c, g, ex};
  return AE_OK;
}
ae_status ae_graph_launch(ae_graph* g) {
  if (!g) return fail(AE_EARG, "null");
  cudaSetDevice(g->c->dev);
  CK(cudaGraphLaunch(g->exec, g->c->stream));
  return AE_OK;
}
ae_status ae_graph_destroy(ae_graph* g) {
  if (!g) return AE_OK;
  cudaSetDevice(g->c->dev);
  cudaGraphExecDestroy(g->exec);
  cudaGraphDestroy(g->graph);
  delete g;
  return AE_OK;
}
}  // extern "C"

// -------------------------------------------------------------------------------------------------
// Streaming form of the host pipeline: the src/pipeline.rs:26-137 + src/pool.rs:43-130 analogue for
// the one place this path has stages.  Stages are the copy engines and the SMs instead of threads;
// the "pool" is a ring of `depth` slots (device input block, device bit block, stream, events), and
// the per-stage report (processed, active time, rate, utilisation: pipeline.rs:93-107) comes from
// CUDA events: a stage's active time is the union of its [start, end] intervals.
// -------------------------------------------------------------------------------------------------
struct PipeSlot {
  cudaStream_t st = nullptr;
  float2* d_in = nullptr;
  uint8_t* d_out = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // before H2D, after H2D, after kernel, after D2H
  uint8_t* host_bits = nullptr;
  bool busy = false;
};
struct ae_pipe {
  ae_chain* ch;
  size_t block_frames;
  std::vector<PipeSlot> slots;
  std::vector<uint8_t*> done;          // retired blocks not yet handed out by ae_pipe_recv (FIFO)
  size_t head = 0, tail = 0;           // next slot to send into / oldest slot in flight
  cudaEvent_t t0 = nullptr;            // time origin of the current report window
  double active_ms[3] = {0, 0, 0}, last_end_ms[3] = {0, 0, 0}, window_end_ms = 0;
  uint64_t processed = 0;
};

namespace {
ae_status pipe_retire(ae_pipe* p, PipeSlot& s) {
  CK(cudaEventSynchronize(s.ev[3]));
  float t[4];
  for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&t[i], p->t0, s.ev[i]));
  for (int k = 0; k < 3; ++k) {
    const double start = std::max((double)t[k], p->last_end_ms[k]), end = t[k + 1];
    if (end > start) p->active_ms[k] += end - start;
    p->last_end_ms[k] = std::max(p->last_end_ms[k], end);
  }
  p->window_end_ms = std::max(p->window_end_ms, (double)t[3]);
  p->processed += 1;
  p->done.push_back(s.host_bits);
  s.busy = false;
  p->tail += 1;
  return AE_OK;
}
}  // namespace

extern "C" {

ae_status ae_pipe_create(ae_chain* ch, size_t block_frames, int depth, ae_pipe** out) {
  if (!ch || !out) return fail(AE_EARG, "null");
  if (!ch->fused) return fail(AE_EARG, "host pipeline needs the fused chain (power-of-two FFT length 256..4096)");
  if (block_frames == 0 || depth < 1 || depth > 16) return fail(AE_EARG, "block_frames >= 1 and 1 <= depth <= 16");
  cudaSetDevice(ch->c->dev);
  TRY(ae_sync());                      // the chain's window/taps uploads are ordered on the context stream
  ae_pipe* p = new ae_pipe;
  p->ch = ch;
  p->block_frames = block_frames;
  p->slots.resize(depth);
  cudaError_t e = cudaEventCreate(&p->t0);
  for (auto& s : p->slots) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s.d_in, block_frames * ch->n * sizeof(float2));
    if (e == cudaSuccess) e = cudaMalloc((void**)&s.d_out, block_frames * ch->n * 2);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&s.ev[i]);
  }
  if (e == cudaSuccess) e = cudaEventRecord(p->t0, p->slots[0].st);
  if (e != cudaSuccess) {
    ae_pipe_destroy(p);
    return fail(e == cudaErrorMemoryAllocation ? AE_EOOM : AE_ECUDA, std::string("ae_pipe_create: ") + cudaGetErrorString(e));
  }
  *out = p;
  return AE_OK;
}

ae_status ae_pipe_destroy(ae_pipe* p) {
  if (!p) return AE_OK;
  cudaSetDevice(p->ch->c->dev);
  for (auto& s : p->slots) {
    if (s.st) cudaStreamSynchronize(s.st);
    for (auto& ev : s.ev) if (ev) cudaEventDestroy(ev);
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.st) cudaStreamDestroy(s.st);
  }
  if (p->t0) cudaEventDestroy(p->t0);
  delete p;
  return AE_OK;
}

ae_status ae_pipe_send(ae_pipe* p, const ae_cf32* host_in, uint8_t* host_bits) {
  if (!p || !host_in || !host_bits) return fail(AE_EARG, "null");
  ae_chain* ch = p->ch;
  cudaSetDevice(ch->c->dev);
  PipeSlot& s = p->slots[p->head % p->slots.size()];
  if (s.busy) TRY(pipe_retire(p, s));  // every slot in flight: wait for the oldest block (Pool::take on an empty pool)
  const size_t ns = p->block_frames * ch->n;
  CK(cudaEventRecord(s.ev[0], s.st));
  CK(cudaMemcpyAsync(s.d_in, host_in, ns * sizeof(float2), cudaMemcpyHostToDevice, s.st));
  CK(cudaEventRecord(s.ev[1], s.st));
  chain_launch(ch, s.d_in, s.d_out, p->block_frames, s.st);
  CKL(1);
  CK(cudaEventRecord(s.ev[2], s.st));
  CK(cudaMemcpyAsync(host_bits, s.d_out, ns * 2, cudaMemcpyDeviceToHost, s.st));
  CK(cudaEventRecord(s.ev[3], s.st));
  s.host_bits = host_bits;
  s.busy = true;
  p->head += 1;
  return AE_OK;
}

ae_status ae_pipe_recv(ae_pipe* p, uint8_t** host_bits_done) {
  if (!p || !host_bits_done) return fail(AE_EARG, "null");
  cudaSetDevice(p->ch->c->dev);
  if (p->done.empty()) {
    if (p->tail == p->head) return fail(AE_EARG, "ae_pipe_recv: nothing in flight");
    TRY(pipe_retire(p, p->slots[p->tail % p->slots.size()]));
  }
  *host_bits_done = p->done.front();
  p->done.erase(p->done.begin());
  return AE_OK;
}

size_t ae_pipe_in_flight(const ae_pipe* p) { return p ? (p->head - p->tail) + p->done.size() : 0; }

ae_status ae_pipe_report(ae_pipe* p, ae_pipe_stage stages[3], int reset) {
  if (!p || !stages) return fail(AE_EARG, "null");
  static const char* names[3] = {"h2d", "fft-fir-demod", "d2h"};
  for (int k = 0; k < 3; ++k) {
    ae_pipe_stage& r = stages[k];
    memset(&r, 0, sizeof(r));
    snprintf(r.name, sizeof(r.name), "%s", names[k]);
    r.processed = p->processed;
    r.active_ms = p->active_ms[k];
    r.elapsed_ms = p->window_end_ms;
    r.per_second = p->window_end_ms > 0 ? p->processed / p->window_end_ms * 1e3 : 0.0;
    r.utilisation_pct = p->window_end_ms > 0 ? 100.0 * p->active_ms[k] / p->window_end_ms : 0.0;
  }
  if (reset) {
    if (p->tail != p->head) return fail(AE_EARG, "ae_pipe_report(reset): blocks still in flight");
    cudaSetDevice(p->ch->c->dev);
    CK(cudaEventRecord(p->t0, p->slots[0].st));
    for (int k = 0; k < 3; ++k) p->active_ms[k] = p->last_end_ms[k] = 0;
    p->window_end_ms = 0;
    p->processed = 0;
  }
  return AE_OK;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// General stage pipeline (src/pipeline.rs:26-137): a CUDA stream per stage instead of a thread, events
// instead of channels.  Every stage's op runs on the calling thread inside ae_pipeline_send and only QUEUES
// work: while it runs, the context stream is the stage's stream.
// -------------------------------------------------------------------------------------------------
struct FlowStage {
  std::string name;
  ae_stage_fn op = nullptr;
  void* user = nullptr;
  cudaStream_t st = nullptr;
  std::vector<cudaEvent_t> ev0, ev1;    // per slot: before / after the stage's work for the item in that slot
  double active_ms = 0, last_end_ms = 0;
};
struct ae_pipeline {
  Ctx* c = nullptr;
  int depth = 0;
  std::vector<FlowStage> stages;
  std::vector<void*> item;              // per slot
  std::vector<char> busy;
  std::vector<void*> done;              // retired items not yet handed out (FIFO)
  size_t head = 0, tail = 0;
  cudaEvent_t t0 = nullptr, fence = nullptr;
  double window_end_ms = 0;
  uint64_t processed = 0;
  bool sent_any = false;
};

namespace {
ae_status flush_all_pending(Ctx* c) {
  std::vector<ae_vec*> snap = c->pending;
  for (ae_vec* v : snap) TRY(flush_vec(v));
  return AE_OK;
}
ae_status flow_retire(ae_pipeline* p, size_t slot) {
  FlowStage& last = p->stages.back();
  CK(cudaEventSynchronize(last.ev1[slot]));
  float t_end = 0;
  for (FlowStage& s : p->stages) {
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, p->t0, s.ev0[slot]));
    CK(cudaEventElapsedTime(&b, p->t0, s.ev1[slot]));
    const double start = std::max((double)a, s.last_end_ms), end = b;
    if (end > start) s.active_ms += end - start;
    s.last_end_ms = std::max(s.last_end_ms, end);
    t_end = std::max(t_end, b);
  }
  p->window_end_ms = std::max(p->window_end_ms, (double)t_end);
  p->processed += 1;
  p->busy[slot] = 0;
  p->done.push_back(p->item[slot]);
  p->tail += 1;
  return AE_OK;
}
}  // namespace

extern "C" {
ae_status ae_pipeline_create(int depth, ae_pipeline** out) {
  if (!out) return fail(AE_EARG, "null");
  if (depth < 1 || depth > 64) return fail(AE_EARG, "ae_pipeline_create: depth must be in 1..64");
  Ctx* c;
  TRY(get_ctx(&c));
  ae_pipeline* p = new ae_pipeline;
  p->c = c;
  p->depth = depth;
  p->item.assign((size_t)depth, nullptr);
  p->busy.assign((size_t)depth, 0);
  if (cudaEventCreate(&p->t0) != cudaSuccess || cudaEventCreateWithFlags(&p->fence, cudaEventDisableTiming) != cudaSuccess) {
    ae_pipeline_destroy(p);
    return fail(AE_ECUDA, "ae_pipeline_create: cudaEventCreate failed");
  }
  *out = p;
  return AE_OK;
}
ae_status ae_pipeline_add_stage(ae_pipeline* p, const char* name, ae_stage_fn op, void* user) {
  if (!p || !op) return fail(AE_EARG, "null");
  if (p->sent_any) return fail(AE_EARG, "ae_pipeline_add_stage: the pipeline is already running");
  cudaSetDevice(p->c->dev);
  FlowStage s;
  s.name = name ? name : "";
  s.op = op;
  s.user = user;
  CK(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
  for (int i = 0; i < p->depth; ++i) {
    cudaEvent_t a = nullptr, b = nullptr;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
      if (a) cudaEventDestroy(a);
      for (cudaEvent_t e : s.ev0) cudaEventDestroy(e);
      for (cudaEvent_t e : s.ev1) cudaEventDestroy(e);
      cudaStreamDestroy(s.st);
      return fail(AE_ECUDA, "ae_pipeline_add_stage: cudaEventCreate failed");
    }
    s.ev0.push_back(a);
    s.ev1.push_back(b);
  }
  p->stages.push_back(std::move(s));
  return AE_OK;
}
ae_status ae_pipeline_send(ae_pipeline* p, void* item) {
  if (!p) return fail(AE_EARG, "null");
  if (p->stages.empty()) return fail(AE_EARG, "ae_pipeline_send: no stages");
  Ctx* c = p->c;
  cudaSetDevice(c->dev);
  const size_t slot = p->head % (size_t)p->depth;
  if (p->busy[slot]) TRY(flow_retire(p, slot));         // `depth` items in flight: wait for the oldest one
  // whatever the caller queued on the context stream so far (recorded VecOps included) happens before the first stage
  TRY(flush_all_pending(c));
  cudaStream_t user_stream = c->stream;
  if (!p->sent_any) {
    CK(cudaEventRecord(p->t0, user_stream));
    p->sent_any = true;
  }
  CK(cudaEventRecord(p->fence, user_stream));
  CK(cudaStreamWaitEvent(p->stages[0].st, p->fence, 0));
  if (p->head == 0) CK(cudaStreamWaitEvent(p->stages[0].st, p->t0, 0));
  ae_status st = AE_OK;
  for (size_t k = 0; k < p->stages.size() && st == AE_OK; ++k) {
    FlowStage& s = p->stages[k];
    cudaError_t e = cudaSuccess;
    if (k > 0) e = cudaStreamWaitEvent(s.st, p->stages[k - 1].ev1[slot], 0);
    if (e == cudaSuccess) e = cudaEventRecord(s.ev0[slot], s.st);
    if (e != cudaSuccess) { st = fail(AE_ECUDA, std::string(cudaGetErrorString(e)) + " in ae_pipeline_send"); break; }
    c->stream = s.st;                                    // the stage's "thread"
    st = s.op(s.user, slot, item);
    if (st == AE_OK) st = flush_all_pending(c);          // VecOps the stage recorded run on ITS stream
    c->stream = user_stream;
    e = cudaEventRecord(s.ev1[slot], s.st);
    if (st == AE_OK && e != cudaSuccess) st = fail(AE_ECUDA, std::string(cudaGetErrorString(e)) + " in ae_pipeline_send");
  }
  if (st != AE_OK) {
    // a failed stage leaves the item half processed: drain and report; the item is not in flight
    for (FlowStage& s : p->stages) cudaStreamSynchronize(s.st);
    return st;
  }
  p->item[slot] = item;
  p->busy[slot] = 1;
  p->head += 1;
  return AE_OK;
}
ae_status ae_pipeline_recv(ae_pipeline* p, void** item_done) {
  if (!p || !item_done) return fail(AE_EARG, "null");
  cudaSetDevice(p->c->dev);
  if (p->done.empty()) {
    if (p->tail == p->head) return fail(AE_EARG, "ae_pipeline_recv: nothing in flight");
    TRY(flow_retire(p, p->tail % (size_t)p->depth));
  }
  *item_done = p->done.front();
  p->done.erase(p->done.begin());
  return AE_OK;
}
size_t ae_pipeline_in_flight(const ae_pipeline* p) { return p ? (p->head - p->tail) + p->done.size() : 0; }
size_t ae_pipeline_stages(const ae_pipeline* p) { return p ? p->stages.size() : 0; }
ae_status ae_pipeline_report(ae_pipeline* p, ae_pipe_stage* stages, size_t n_stages, int reset) {
  if (!p || (!stages && n_stages)) return fail(AE_EARG, "null");
  if (n_stages < p->stages.size()) return fail(AE_ELEN, "ae_pipeline_report: room for every stage is required");
  for (size_t k = 0; k < p->stages.size(); ++k) {
    ae_pipe_stage& r = stages[k];
    memset(&r, 0, sizeof(r));
    snprintf(r.name, sizeof(r.name), "%s", p->stages[k].name.c_str());
    r.processed = p->processed;
    r.active_ms = p->stages[k].active_ms;
    r.elapsed_ms = p->window_end_ms;
    r.per_second = p->window_end_ms > 0 ? p->processed / p->window_end_ms * 1e3 : 0.0;
    r.utilisation_pct = p->window_end_ms > 0 ? 100.0 * p->stages[k].active_ms / p->window_end_ms : 0.0;
  }
  if (reset) {
    if (p->tail != p->head) return fail(AE_EARG, "ae_pipeline_report(reset): items still in flight");
    cudaSetDevice(p->c->dev);
    for (FlowStage& s : p->stages) {
      CK(cudaStreamSynchronize(s.st));
      s.active_ms = s.last_end_ms = 0;
    }
    CK(cudaEventRecord(p->t0, p->c->stream));
    CK(cudaStreamWaitEvent(p->stages.empty() ? p->c->stream : p->stages[0].st, p->t0, 0));
    p->window_end_ms = 0;
    p->processed = 0;
  }
  return AE_OK;
}
ae_status ae_pipeline_destroy(ae_pipeline* p) {
  if (!p) return AE_OK;
  cudaSetDevice(p->c->dev);
  for (FlowStage& s : p->stages) {
    if (s.st) cudaStreamSynchronize(s.st);
    for (cudaEvent_t e : s.ev0) cudaEventDestroy(e);
    for (cudaEvent_t e : s.ev1) cudaEventDestroy(e);
    if (s.st) cudaStreamDestroy(s.st);
  }
  if (p->t0) cudaEventDestroy(p->t0);
  if (p->fence) cudaEventDestroy(p->fence);
  delete p;
  return AE_OK;
}
}  // extern "C"

extern "C" {
ae_status ae_ofdm_chain(size_t fft_len, size_t frames, uint64_t first_frame_id, float noise_power, uint64_t noise_seed,
                        int compat, ae_bits* tx_bits, ae_bits* rx_bits, ae_stats* d) {
  if (!ofdm_supported(fft_len)) return fail(AE_EARG, "ofdm chain supports power-of-two FFT lengths 512..4096");
  Ctx* c;
  TRY(get_ctx(&c));
  float2* tw;
  TRY(get_thread_twiddles(c, fft_len, &tw));
  const size_t nbits = 2 * fft_len * frames;
  if (tx_bits) { TRY(bits_reserve(tx_bits, nbits)); tx_bits->len = nbits; }
  if (rx_bits) { TRY(bits_reserve(rx_bits, nbits)); rx_bits->len = nbits; }
  launch_ofdm_chain(fft_len, frames, first_frame_id, sqrtf(noise_power), compat == AE_COMPAT_REFERENCE ? 1 : 0, noise_seed, tw,
                    compat, tx_bits ? bptr(tx_bits) : nullptr, rx_bits ? bptr(rx_bits) : nullptr, d, c->stream);
  if (frames) CKL(1);
  return AE_OK;
}

// =================================================================================================
// spectrogram core and correlator
// =================================================================================================
}  // extern "C"

struct ae_f32 {
  Ctx* c;
  float* p;
  size_t len, cap;
};

extern "C" {

ae_status ae_f32_alloc(size_t len, ae_f32** out) {
  if (!out) return fail(AE_EARG, "null");
  Ctx* c;
  TRY(get_ctx(&c));
  void* p = nullptr;
  TRY(dev_alloc(c, len * sizeof(float), &p));
  ae_f32* v = new ae_f32;
  v->c = c; v->p = (float*)p; v->len = len; v->cap = len;
  *out = v;
  return AE_OK;
}
ae_status ae_f32_free(ae_f32* v) {
  if (!v) return AE_OK;
  cudaSetDevice(v->c->dev);
  dev_free(v->c, v->p);
  delete v;
  return AE_OK;
}
size_t ae_f32_len(const ae_f32* v) { return v ? v->len : 0; }
ae_status ae_f32_device_ptr(ae_f32* v, void** ptr) {
  if (!v || !ptr) return fail(AE_EARG, "null");
  *ptr = v->p;
  return AE_OK;
}
ae_status ae_f32_download(ae_f32* v, float* host, size_t n) {
  if (!v || (!host && n)) return fail(AE_EARG, "null");
  if (n != v->len) return fail(AE_ELEN, "Vectors must have same length");
  cudaSetDevice(v->c->dev);
  if (n) CK(cudaMemcpyAsync(host, v->p, n * sizeof(float), cudaMemcpyDeviceToHost, v->c->stream));
  return sync_ctx(v->c);
}

ae_status ae_f32_stats(ae_f32* v, ae_vecstats* host_out) {
  if (!v) return fail(AE_EARG, "null");
  return vecstats_impl(v->c, v->p, v->len, false, host_out);
}

ae_status ae_spectrogram(ae_fft* f, ae_vec* symbols, ae_f32* levels, int use_db) {
  if (!f || !symbols || !levels) return fail(AE_EARG, "null");
  Ctx* c = f->c;
  cudaSetDevice(c->dev);
  const size_t n = f->len;
  const size_t chunks = (symbols->len + n - 1) / n;   // padded.len() % fft_len == 0 (src/util/plot.rs:52-58)
  const size_t total = chunks * n;
  if (levels->cap < total) {
    dev_free(c, levels->p);
    levels->p = nullptr; levels->cap = 0;
    void* p;
    TRY(dev_alloc(c, total * sizeof(float), &p));
    levels->p = (float*)p; levels->cap = total;
  }
  levels->len = total;
  if (total == 0) return AE_OK;
  TRY(before_read(symbols));
  const bool inverse = (f->compat == AE_COMPAT_REFERENCE);  // vec_rfft = Fft::ifwd
  const float s = scale_factor(AE_SCALE_SN, n, 1.0f);
  if (f->pow2 && spectral_supported(n)) {
    launch_spectrogram(vptr(symbols), symbols->len, levels->p, n, chunks, f->tw, inverse, s, use_db, c->stream);
    CKL(1);
    return AE_OK;
  }
  // any other length: pad, transform with the stand-alone kernels, then mirror + level
  void* tmp;
  TRY(dev_alloc(c, total * sizeof(float2), &tmp));
  CK(cudaMemsetAsync(tmp, 0, total * sizeof(float2), c->stream));
  CK(cudaMemcpyAsync(tmp, vptr(symbols), symbols->len * sizeof(float2), cudaMemcpyDeviceToDevice, c->stream));
  ae_status st = fft_run(f, AE_FFT_FWD, (const float2*)tmp, (float2*)tmp, AE_SCALE_SN, 1.0f, chunks);
  if (st == AE_OK) {
    launch_levels((const float2*)tmp, levels->p, total, n, use_db, c->stream);
    g_launches += 1;
  }
  dev_free(c, tmp);
  return st;
}

ae_status ae_correlate(ae_fft* f, ae_vec* inout, ae_vec* sig, int scale_kind, float x, size_t howmany) {
  if (!f || !inout || !sig) return fail(AE_EARG, "null");
  if (scale_kind < 0 || scale_kind > 3) return fail(AE_EARG, "bad scale kind");
  if (inout->len != f->len * howmany) return fail(AE_ELEN, "Input and FFT must be the same length");  // src/fft.rs:185-189
  if (sig->len != f->len) return fail(AE_ELEN, "Vectors must have same length");                       // src/vecops.rs:100-104
  Ctx* c = f->c;
  cudaSetDevice(c->dev);
  if (howmany == 0) return AE_OK;
  TRY(before_write(inout));
  TRY(before_read(sig));
  const size_t n = f->len;
  const float s = scale_factor(scale_kind, n, x);
  static const char* corr_v1 = getenv("AE_CORR_V1");   // developer switch: the two-exchange kernel for 1024 points too
  if (n == 1024 && !corr_v1) {
    float2* x2tw;
    TRY(get_x2_twiddles(c, &x2tw));
    launch_correlate_x2(vptr(inout), vptr(sig), howmany, x2tw, f->compat == AE_COMPAT_REFERENCE, s, scale_kind != AE_SCALE_NONE, c->stream);
    CKL(1);
    return AE_OK;
  }
  if (f->pow2 && spectral_supported(n)) {
    launch_correlate(vptr(inout), vptr(sig), n, howmany, f->tw, f->compat == AE_COMPAT_REFERENCE, s, scale_kind != AE_SCALE_NONE,
                     c->stream);
    CKL(1);
    return AE_OK;
  }
  // composition of the stand-alone kernels, frame by frame for the multiply
  TRY(fft_run(f, AE_FFT_FWD, vptr(inout), vptr(inout), scale_kind, x, howmany));
  for (size_t fr = 0; fr < howmany; ++fr) {
    TapeParams tp;
    std::memset(&tp, 0, sizeof(tp));
    tp.n_ops = 1; tp.load_self = 1;
    tp.e[0].op = OP_MUL; tp.e[0].operand = vptr(sig);
    launch_vecops(vptr(inout) + fr * n, n, tp, false, c->sm_count, c->stream);
    CKL(1);
  }
  return fft_run(f, AE_FFT_BWD, vptr(inout), vptr(inout), scale_kind, x, howmany);
}

ae_status ae_awgn_next_host(ae_awgn* g, ae_cf32* host_out, size_t n) {
  if (!g || (!host_out && n)) return fail(AE_EARG, "null");
  if (n == 0) return AE_OK;
  ae_vec* v;
  TRY(ae_vec_alloc(0, n, &v));
  ae_status st = ae_awgn_fill(g, v);                       // NoiseIter::next == Awgn::next (src/noise.rs:81-83)
  if (st == AE_OK) st = ae_vec_download(v, host_out, n);
  ae_vec_free(v);
  return st;
}

ae_status ae_vec_read_raw(const char* path, ae_vec** out) {
  if (!path || !out) return fail(AE_EARG, "null");
  Ctx* c;
  TRY(get_ctx(&c));
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return fail(AE_EARG, std::string("cannot open ") + path);
  std::fseek(fp, 0, SEEK_END);
  const long long bytes = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  if (bytes < 0 || bytes % (long long)sizeof(ae_cf32)) {   // count_structs_in_file (src/util/file.rs:12-25)
    std::fclose(fp);
    return fail(AE_EARG, "File does not contain an integer number of the requested struct");
  }
  const size_t n = (size_t)bytes / sizeof(ae_cf32);
  ae_vec* v;
  ae_status st = ae_vec_alloc(n, n, &v);
  if (st != AE_OK) { std::fclose(fp); return st; }
  // two pinned staging buffers: read chunk i+1 from disk while chunk i is copied to the device
  const size_t chunk = (size_t)8 << 20;  // cf32 per chunk (64 MiB)
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) {
    if (cudaHostAlloc(&stage[i], std::min(chunk, std::max<size_t>(n, 1)) * sizeof(ae_cf32), cudaHostAllocDefault) != cudaSuccess) st = fail(AE_EOOM, "pinned staging");
    cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  }
  size_t done = 0;
  int slot = 0;
  while (st == AE_OK && done < n) {
    const size_t k = std::min(chunk, n - done);
    cudaEventSynchronize(ev[slot]);  // the previous copy out of this buffer has finished
    if (std::fread(stage[slot], sizeof(ae_cf32), k, fp) != k) { st = fail(AE_EARG, "failed to fill whole buffer"); break; }
    if (cudaMemcpyAsync(vptr(v) + done, stage[slot], k * sizeof(ae_cf32), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { st = fail(AE_ECUDA, "H2D"); break; }
    cudaEventRecord(ev[slot], c->stream);
    done += k;
    slot ^= 1;
  }
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; ++i) { cudaFreeHost(stage[i]); cudaEventDestroy(ev[i]); }
  std::fclose(fp);
  if (st != AE_OK) { ae_vec_free(v); return st; }
  *out = v;
  return AE_OK;
}

ae_status ae_vec_write_raw(ae_vec* v, const char* path) {
  if (!v || !path) return fail(AE_EARG, "null");
  std::vector<ae_cf32> h(v->len);
  TRY(ae_vec_download(v, h.data(), h.size()));
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return fail(AE_EARG, std::string("cannot open ") + path);
  const size_t w = std::fwrite(h.data(), sizeof(ae_cf32), h.size(), fp);
  std::fclose(fp);
  if (w != h.size()) return fail(AE_EARG, "short write");
  return AE_OK;
}

}  // extern "C"
