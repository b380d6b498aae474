// Small CUDA utilities shared by the pgmm kernels: loud failure, device buffers, stream-ordered helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <map>
#include <mutex>
#include <vector>

// CUDA failures abort with a message: the reference's C has no error channel either (SURVEY 8b "Errors").
#define PGMM_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      fprintf(stderr, "[pgmm_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__,    \
              __LINE__, cudaGetErrorString(e_));                                                     \
      abort();                                                                                       \
    }                                                                                                \
  } while (0)

#define PGMM_FATAL(...)                        \
  do {                                         \
    fprintf(stderr, "[pgmm_b200] fatal: ");    \
    fprintf(stderr, __VA_ARGS__);              \
    fprintf(stderr, "\n");                     \
    abort();                                   \
  } while (0)

namespace pgmm {

// bytes moved across the bus, counted where they are issued (read by pgmm_get_stats)
inline uint64_t &h2d_bytes() {
  static uint64_t v = 0;
  return v;
}
inline uint64_t &d2h_bytes() {
  static uint64_t v = 0;
  return v;
}
inline cudaError_t counted_memcpy_async(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t st) {
  if (kind == cudaMemcpyHostToDevice) __atomic_fetch_add(&h2d_bytes(), (uint64_t)n, __ATOMIC_RELAXED);
  else if (kind == cudaMemcpyDeviceToHost) __atomic_fetch_add(&d2h_bytes(), (uint64_t)n, __ATOMIC_RELAXED);
  return cudaMemcpyAsync(dst, src, n, kind, st);
}

// Process-wide cache of device allocations: freed blocks are kept per size class and handed out again, so that the
// steady state of a build (index after index, round after round) never calls cudaMalloc/cudaFree.
class DevicePool {
 public:
  static DevicePool &get() {
    static DevicePool *p = new DevicePool;  // intentionally leaked: outlives every static DevBuf
    return *p;
  }
  // size classes: 8 steps per power of two
  static size_t size_class(size_t bytes) {
    if (bytes < 4096) bytes = 4096;
    size_t p2 = 4096;
    while (p2 < bytes) p2 <<= 1;
    const size_t step = p2 >> 4;  // p2/2 .. p2 in 8 steps
    const size_t base = p2 >> 1;
    return base + (bytes - base + step - 1) / step * step;
  }
  void *alloc(size_t cls) {
    {
      std::lock_guard<std::mutex> g(mu_);
      auto it = free_.find(cls);
      if (it != free_.end() && !it->second.empty()) {
        void *p = it->second.back();
        it->second.pop_back();
        return p;
      }
    }
    void *p = nullptr;
    __atomic_fetch_add(&misses(), 1, __ATOMIC_RELAXED);
    cudaError_t e = cudaMalloc(&p, cls);
    if (e != cudaSuccess) {  // give cached blocks back to the driver and retry once
      trim();
      e = cudaMalloc(&p, cls);
    }
    if (e != cudaSuccess) {
      size_t fr = 0, tot = 0;
      cudaMemGetInfo(&fr, &tot);
      PGMM_FATAL("cudaMalloc of %zu bytes failed: %s (%zu of %zu bytes free on the device; the library's budget is set by "
                 "PGMM_DP_ARENA_GB, PGMM_CONTEXTS and what the caller keeps resident)", cls, cudaGetErrorString(e), fr, tot);
    }
    return p;
  }
  void release(void *p, size_t cls) {
    std::lock_guard<std::mutex> g(mu_);
    free_[cls].push_back(p);
  }
  static uint64_t &misses() {  // cudaMalloc calls so far (each one synchronises the device)
    static uint64_t v = 0;
    return v;
  }
  void trim() {
    std::lock_guard<std::mutex> g(mu_);
    for (auto &kv : free_)
      for (void *p : kv.second) cudaFree(p);
    free_.clear();
  }

 private:
  std::mutex mu_;
  std::map<size_t, std::vector<void *>> free_;
};

// Grow-only device buffer backed by the pool.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;   // elements
  size_t cls = 0;   // bytes of the pool block
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) {
      if (exact_) cudaFree(p);
      else DevicePool::get().release(p, cls);
    }
    p = nullptr;
    cap = 0, cls = 0, exact_ = false;
  }
  bool exact_ = false;  // block came straight from cudaMalloc (ensure_exact), not from the pool
  // for the traceback arenas: exact size (2 MB granularity), allocated from and returned to the DRIVER -- a cached old
  // arena of another size would be dead weight in the pool.  Growing costs a cudaFree + cudaMalloc (both synchronise the
  // device), which only happens while the arenas find their steady-state size.
  T *ensure_exact(size_t n) {
    if (n > cap) {
      release();
      const size_t bytes = (n * sizeof(T) + (2u << 20) - 1) / (2u << 20) * (2u << 20);
      void *q = nullptr;
      __atomic_fetch_add(&DevicePool::misses(), 1, __ATOMIC_RELAXED);
      cudaError_t e = cudaMalloc(&q, bytes);
      if (e != cudaSuccess) {
        DevicePool::get().trim();
        e = cudaMalloc(&q, bytes);
      }
      if (e != cudaSuccess) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        PGMM_FATAL("cudaMalloc of a %zu-byte traceback arena failed: %s (%zu of %zu bytes free on the device; lower PGMM_DP_ARENA_GB "
                   "or PGMM_CONTEXTS)", bytes, cudaGetErrorString(e), fr, tot);
      }
      p = (T *)q, cap = bytes / sizeof(T), cls = 0, exact_ = true;
    }
    return p;
  }
  T *ensure(size_t n) {
    if (n > cap) {
      // The old block goes back to the process-wide pool, where another context may pick it up at once: work that this
      // context has queued on it (e.g. the previous of two back-to-back CUB passes sharing one scratch buffer) must be
      // done first.  Growth stops after the first few rounds (25 % headroom), so the device-wide wait is rare.
      if (p) cudaDeviceSynchronize();
      release();
      cls = DevicePool::size_class((n + n / 4 + 64) * sizeof(T));  // 25 % headroom: rounds differ slightly in size
      p = (T *)DevicePool::get().alloc(cls);
      cap = cls / sizeof(T);
    }
    return p;
  }
};

// Grow-only pinned host buffer.
template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  PinBuf() = default;
  PinBuf(const PinBuf &) = delete;
  PinBuf &operator=(const PinBuf &) = delete;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  T *ensure(size_t n) {
    if (n > cap) {
      if (p) cudaFreeHost(p);
      size_t want = n + n / 2 + 64;
      PGMM_CUDA(cudaMallocHost((void **)&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
};

// The one device this process computes on (one process per GPU).  Fixed by pgmm_set_device or by the first call;
// every entry point re-selects it, because a host thread that never touched CUDA starts on device 0.
inline int &bound_device() {
  static int d = -1;
  return d;
}
// Require a usable CUDA device; the library has no CPU path.
inline void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    PGMM_FATAL("no CUDA device available (%s); libpgmm_b200 has no CPU fallback",
               e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  int &d = bound_device();
  if (d < 0) PGMM_CUDA(cudaGetDevice(&d));
  else PGMM_CUDA(cudaSetDevice(d));
}

}  // namespace pgmm

namespace pgmm {
// Waiting for a stream must not burn a host core (the cores are needed for chaining other rounds): block on an event
// created with cudaEventBlockingSync instead of spinning in cudaStreamSynchronize.
inline cudaError_t blocking_stream_sync(cudaStream_t st) {
  thread_local cudaEvent_t ev = nullptr;
  thread_local int ev_dev = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (ev == nullptr || ev_dev != dev) {
    e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
    ev_dev = dev;
  }
  e = cudaEventRecord(ev, st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev);
}
}  // namespace pgmm
#define cudaStreamSynchronize(st) pgmm::blocking_stream_sync((st))

// every copy in the library goes through the counter
#define cudaMemcpyAsync(dst, src, n, kind, st) pgmm::counted_memcpy_async((dst), (src), (n), (kind), (st))

// ---- optional CTA trace (PGMM_CTA_TRACE / pgmm_cta_trace_begin): every traced CTA leaves (kernel, block, SM, start, end
// of its main loop, end) in a device buffer, so that what actually overlaps on the GPU while many rounds are in flight
// can be reconstructed without a timeline profiler.  Each translation unit has its own copy of the three device
// variables; its attach function (trace_attach_*) points them at the shared buffer.
struct PgmmCtaTraceRec {
  unsigned long long t0, t1, t2;  // %globaltimer (ns): start, end of the main loop, end
  unsigned kernel, block, smid, aux;
};
#ifdef __CUDACC__
namespace pgmm {
namespace trace {
static __device__ PgmmCtaTraceRec *d_buf;
static __device__ unsigned long long *d_cnt;
static __device__ unsigned long long d_cap;
__device__ __forceinline__ unsigned long long now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long begin() { return d_buf ? now() : 0ull; }
__device__ __forceinline__ void emit(unsigned kernel, unsigned long long t0, unsigned long long t1, unsigned aux) {
  if (!d_buf) return;
  const unsigned long long idx = atomicAdd(d_cnt, 1ull);
  if (idx >= d_cap) return;
  unsigned sm;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
  PgmmCtaTraceRec r;
  r.t0 = t0, r.t1 = t1, r.t2 = now(), r.kernel = kernel, r.block = blockIdx.x, r.smid = sm, r.aux = aux;
  d_buf[idx] = r;
}
inline void attach(PgmmCtaTraceRec *buf, unsigned long long *cnt, unsigned long long cap) {
  PGMM_CUDA(cudaMemcpyToSymbol(d_buf, &buf, sizeof(buf)));
  PGMM_CUDA(cudaMemcpyToSymbol(d_cnt, &cnt, sizeof(cnt)));
  PGMM_CUDA(cudaMemcpyToSymbol(d_cap, &cap, sizeof(cap)));
}
}  // namespace trace
void trace_attach_ksw(PgmmCtaTraceRec *buf, unsigned long long *cnt, unsigned long long cap);
void trace_attach_chain(PgmmCtaTraceRec *buf, unsigned long long *cnt, unsigned long long cap);
}  // namespace pgmm
#endif
