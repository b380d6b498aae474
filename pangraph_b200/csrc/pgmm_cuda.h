// Small CUDA utilities shared by the pgmm kernels: loud failure, device buffers, stream-ordered helpers.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
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

// Grow-only device buffer (cudaMallocAsync-free: plain cudaMalloc, reused across calls).
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  T *ensure(size_t n) {
    if (n > cap) {
      release();
      size_t want = n + n / 4 + 64;
      PGMM_CUDA(cudaMalloc((void **)&p, want * sizeof(T)));
      cap = want;
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
      size_t want = n + n / 4 + 64;
      PGMM_CUDA(cudaMallocHost((void **)&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
};

// Require a usable CUDA device; the library has no CPU path.
inline void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    PGMM_FATAL("no CUDA device available (%s); libpgmm_b200 has no CPU fallback",
               e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
}

}  // namespace pgmm

// every copy in the library goes through the counter
#define cudaMemcpyAsync(dst, src, n, kind, st) pgmm::counted_memcpy_async((dst), (src), (n), (kind), (st))
