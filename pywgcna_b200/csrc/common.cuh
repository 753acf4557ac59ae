// Shared host/device helpers of libb200wgcna (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/b200wgcna.h"

namespace b200w {

// ---- error plumbing ---------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const char* fmt, ...);
int ensure_device(int device);  // cudaSetDevice + sm_100 check; 0 or negative code
extern std::atomic<long long> g_launches;
extern double g_last_kernel_ms;
int resolve_precision(int precision);  // B200W_PREC_AUTO -> the default contraction arithmetic
inline bool prec_is_i8(int precision) { return precision == B200W_PREC_I8 || precision == B200W_PREC_I8_FAST; }
// lowest digit-pair weight kept by the int8 contraction (tom_i8.cu) for a resolved precision
inline int prec_w_min(int precision) { return precision == B200W_PREC_I8_FAST ? B200W_I8_SLICES : B200W_I8_SLICES - 1; }

#define B200W_CUDA(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return b200w::fail(_e == cudaErrorMemoryAllocation ? B200W_E_NOMEM : B200W_E_CUDA,    \
                         "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                         __LINE__);                                                         \
  } while (0)

#define B200W_LAUNCHED()                                                                    \
  do {                                                                                      \
    b200w::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    B200W_CUDA(cudaGetLastError());                                                         \
  } while (0)

// stream-ordered scratch buffer, freed on the same stream when it goes out of scope
struct Scratch {
  void* p = nullptr;
  cudaStream_t s = nullptr;
  cudaError_t alloc(size_t bytes, cudaStream_t stream) {
    s = stream;
    return cudaMallocAsync(&p, bytes ? bytes : 16, stream);
  }
  ~Scratch() {
    if (p) cudaFreeAsync(p, s);
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

static inline int64_t round_up(int64_t x, int64_t q) { return (x + q - 1) / q * q; }

// ---- device helpers ---------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// The three similarity transforms, wgcna.py:884-889 and :1102-1107, on a clipped correlation.
__device__ __forceinline__ double net_transform(double c, int net_type) {
  if (net_type == B200W_NET_UNSIGNED) return fabs(c);
  if (net_type == B200W_NET_SIGNED) return (1.0 + c) / 2.0;
  return c < 0.0 ? 0.0 : c;
}

// s ** step as numpy evaluates it for the power chain (wgcna.py:906-911): small integer steps by
// multiplication (numpy's own fast path for 1 and 2), anything else through pow().
__device__ __forceinline__ double step_pow(double s, double step, int istep) {
  switch (istep) {
    case 1: return s;
    case 2: return s * s;
    case 3: return s * s * s;
    case 4: { double q = s * s; return q * q; }
    default: return pow(s, step);
  }
}
#endif

}  // namespace b200w
