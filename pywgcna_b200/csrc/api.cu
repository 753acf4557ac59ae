// Host-pointer entry points of the C ABI (include/b200wgcna.h) and the error plumbing.
// These are what pywgcna_b200/wgcna.py binds behind WGCNA.pickSoftThreshold / adjacency /
// TOMsimilarity: H2D of the caller's float64 buffers, the dev_* kernels, D2H of the result.
#include <stdarg.h>
#include <string.h>
#include <sys/mman.h>

#include <chrono>
#include <map>
#include <string>
#include <mutex>
#include <thread>
#include <vector>

#include "tom_common.cuh"

int b200w_tom_compute_i8(const double* dA_rows, const void* d_operand, int64_t op_rows, const double* d_scale,
                         const double* d_k, int64_t n, int64_t row0, int64_t nrows, int tom_type, int tom_denom,
                         int out_mode, double* d_out, cudaStream_t st, int world, int rank, int64_t per,
                         double* const* peer_out, b200w::TomProgress* progress, int w_min,
                         const b200w::TomColumnOperand* col, const unsigned int* d_plane_flags = nullptr,
                         unsigned int plane_epoch = 0);

namespace b200w {

static thread_local std::string t_error;
std::atomic<long long> g_launches{0};
double g_last_kernel_ms = -1.0;

void set_error(const std::string& msg) { t_error = msg; }

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_error = buf;
  return code;
}

int ensure_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(B200W_E_NO_DEVICE, "no CUDA device available (%s); libb200wgcna has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= count)
    return fail(B200W_E_INVALID, "device %d out of range (%d visible)", device, count);
  static std::mutex mu;
  static std::vector<int> major_of;
  {
    std::lock_guard<std::mutex> lk(mu);
    if ((int)major_of.size() != count) major_of.assign(count, -1);
    if (major_of[device] < 0) {
      int mj = 0;
      if (cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess)
        return fail(B200W_E_CUDA, "cannot query device %d", device);
      major_of[device] = mj;
      // keep stream-ordered scratch (cudaMallocAsync in the dev_* calls) in the pool across
      // synchronisations: with the default release threshold of 0 every sync hands the memory back to
      // the driver and the next call pays for mapping it again (bimodal 53 / 90 ms bench steps)
      cudaMemPool_t pool;
      if (!getenv("B200W_POOL_RELEASE") && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      cudaGetLastError();
    }
    if (major_of[device] != 10)
      return fail(B200W_E_NO_DEVICE, "device %d is sm_%dx; this library is built for sm_100a only",
                  device, major_of[device]);
  }
  B200W_CUDA(cudaSetDevice(device));
  return 0;
}

// Device workspace of the host entry points.  cudaMalloc / cudaFree of the multi-GB buffers (adjacency,
// operand, result) cost tens to hundreds of ms per call and synchronise the device, so freed buffers are
// kept per device and handed out again (smallest cached buffer of at least the requested size and at most
// 25 % larger).  b200w_dev_trim() gives everything back; at most 70 % of the device memory is cached.
struct DevCache {
  std::mutex mu;
  struct Item { int dev; size_t bytes; void* p; };
  std::vector<Item> free_;
  std::map<void*, std::pair<int, size_t>> live;
  size_t cached = 0;
  cudaError_t get(size_t bytes, void** out) {
    int dev = 0;
    cudaGetDevice(&dev);
    {
      std::lock_guard<std::mutex> lk(mu);
      int best = -1;
      for (int i = 0; i < (int)free_.size(); ++i)
        if (free_[i].dev == dev && free_[i].bytes >= bytes && free_[i].bytes <= bytes + bytes / 4 + 4096 &&
            (best < 0 || free_[i].bytes < free_[best].bytes))
          best = i;
      if (best >= 0) {
        *out = free_[best].p;
        live[*out] = {dev, free_[best].bytes};
        cached -= free_[best].bytes;
        free_.erase(free_.begin() + best);
        return cudaSuccess;
      }
    }
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {  // make room and retry once
      cudaGetLastError();
      trim();
      e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) {
      std::lock_guard<std::mutex> lk(mu);
      live[*out] = {dev, bytes};
    }
    return e;
  }
  void put(void* p) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = live.find(p);
    if (it == live.end()) {
      cudaFree(p);
      return;
    }
    const int dev = it->second.first;
    const size_t bytes = it->second.second;
    live.erase(it);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (bytes < ((size_t)1 << 20) || cached + bytes > total_b / 10 * 7) {
      cudaFree(p);
    } else {
      free_.push_back({dev, bytes, p});
      cached += bytes;
    }
  }
  void trim() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& f : free_) cudaFree(f.p);
    free_.clear();
    cached = 0;
  }
};
static DevCache g_dev_cache;

// RAII device buffer + stream for the host entry points
struct DevBuf {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) { return g_dev_cache.get(bytes ? bytes : 16, &p); }
  ~DevBuf() {
    if (p) g_dev_cache.put(p);
  }
  template <class T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};
struct Stream {
  cudaStream_t s = nullptr;
  cudaError_t create() { return cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking); }
  ~Stream() {
    if (s) {
      cudaStreamSynchronize(s);  // an early error return must not leave copies in flight on recycled buffers
      cudaStreamDestroy(s);
    }
  }
};
// Declared AFTER the buffers of a host entry point, so it runs before they are released: on an early
// `return e` the work already queued on the call's stream is drained before DevBuf hands the memory back
// to the cache (where the next call could reuse it on another stream) and HostPin unregisters the
// caller's pages.  On the success path the stream is idle already and this costs nothing.
struct StreamDrain {
  cudaStream_t s;
  ~StreamDrain() {
    if (cudaStreamSynchronize(s) != cudaSuccess) cudaGetLastError();
  }
};

// ---- host memory -------------------------------------------------------------------------
// (1) HostPin: page-locks a caller's (pageable) buffer for the duration of one call, so that the n x n
//     transfers run as DMA at PCIe speed (measured on the B200 box: 3.2 GB pageable H2D 280 ms, fresh
//     np.empty D2H 790 ms; cudaHostRegister 63 ms + 58 ms DMA + 47 ms unregister).
// (2) a pool of page-locked buffers for the n x n RESULTS (b200w_host_alloc / b200w_host_free): first-touch
//     page faults of a fresh 3.2 GB numpy array cost ~0.5 s and cudaHostAlloc 1.75 s, a recycled pinned
//     buffer costs nothing and is written by DMA at 52 GB/s.
// B200W_TRACE=1: per-phase wall times of the host entry points on stderr (every mark synchronises the
// stream, so the phases are serialised while tracing)
struct Trace {
  const char* name;
  cudaStream_t s;
  bool on;
  std::chrono::steady_clock::time_point t0;
  std::string line;
  Trace(const char* name_, cudaStream_t s_) : name(name_), s(s_) {
    static const bool enabled = getenv("B200W_TRACE") != nullptr;
    on = enabled;
    if (on) t0 = std::chrono::steady_clock::now();
  }
  void mark(const char* what) {
    if (!on) return;
    if (s) cudaStreamSynchronize(s);
    const auto t1 = std::chrono::steady_clock::now();
    char buf[96];
    snprintf(buf, sizeof buf, "  %s %.2f ms", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    line += buf;
    t0 = t1;
  }
  ~Trace() {
    if (on) fprintf(stderr, "[b200w] %s:%s\n", name, line.c_str());
  }
};

struct HostPin {
  void* p = nullptr;
  void pin(const void* ptr, size_t bytes) {
    if (!ptr || bytes < ((size_t)32 << 20)) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return;
    cudaGetLastError();
    if (cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess)
      p = const_cast<void*>(ptr);
    else
      cudaGetLastError();  // fall back to a staged copy
  }
  ~HostPin() {
    if (p) cudaHostUnregister(p);
  }
};

struct HostPool {
  std::mutex mu;
  std::map<void*, size_t> live;                 // handed out: ptr -> bytes
  std::map<void*, bool> registered;             // ptr -> page-locked?
  std::vector<std::pair<size_t, void*>> free_;  // cached
  size_t cached = 0;
  size_t cap() {
    const char* e = getenv("B200WGCNA_PINNED_POOL_MB");
    return e ? (size_t)atoll(e) << 20 : (size_t)96 << 30;
  }
  void* alloc(size_t bytes) {
    const size_t sz = (size_t)round_up((int64_t)(bytes ? bytes : 1), (int64_t)2 << 20);
    {
      std::lock_guard<std::mutex> lk(mu);
      for (size_t i = 0; i < free_.size(); ++i)
        if (free_[i].first == sz) {
          void* p = free_[i].second;
          free_.erase(free_.begin() + i);
          cached -= sz;
          live[p] = sz;
          return p;
        }
    }
    void* p = mmap(nullptr, sz, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    madvise(p, sz, MADV_HUGEPAGE);
    // first touch in parallel (page faults are the cost, not the stores)
    const unsigned nt = sz >= ((size_t)256 << 20) ? 8u : 1u;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([=] {
        char* b = (char*)p + sz / nt * t;
        char* e = (t + 1 == nt) ? (char*)p + sz : b + sz / nt;
        for (char* q = b; q < e; q += 4096) *q = 0;
      });
    for (auto& t : th) t.join();
    bool reg = cudaHostRegister(p, sz, cudaHostRegisterPortable) == cudaSuccess;
    if (!reg) cudaGetLastError();
    std::lock_guard<std::mutex> lk(mu);
    live[p] = sz;
    registered[p] = reg;
    return p;
  }
  void release(void* p, size_t sz) {
    if (registered[p]) cudaHostUnregister(p);
    registered.erase(p);
    munmap(p, sz);
  }
  void free(void* p) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = live.find(p);
    if (it == live.end()) return;
    const size_t sz = it->second;
    live.erase(it);
    if (cached + sz > cap()) {
      release(p, sz);
    } else {
      free_.push_back({sz, p});
      cached += sz;
    }
  }
  void trim() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& f : free_) release(f.second, f.first);
    free_.clear();
    cached = 0;
  }
};
static HostPool g_pool;

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_dev_malloc(int64_t bytes, int device, void** out) {
  if (!out || bytes < 0) return fail(B200W_E_INVALID, "dev_malloc: bad args");
  if (int e = ensure_device(device)) return e;
  B200W_CUDA(cudaMalloc(out, bytes ? (size_t)bytes : 16));
  return 0;
}
extern "C" int b200w_dev_free(void* p, int device) {
  if (int e = ensure_device(device)) return e;
  B200W_CUDA(cudaFree(p));
  return 0;
}
extern "C" int b200w_ipc_export(void* p, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!p || !handle64) return fail(B200W_E_INVALID, "ipc_export: bad args");
  cudaIpcMemHandle_t h;
  B200W_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle64, &h, 64);
  return 0;
}
extern "C" int b200w_ipc_open(const unsigned char* handle64, int device, void** out) {
  if (!handle64 || !out) return fail(B200W_E_INVALID, "ipc_open: bad args");
  if (int e = ensure_device(device)) return e;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  B200W_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int b200w_ipc_close(void* p, int device) {
  if (int e = ensure_device(device)) return e;
  B200W_CUDA(cudaIpcCloseMemHandle(p));
  return 0;
}

extern "C" int b200w_dev_download(void* host_dst, const void* dev_src, int64_t bytes, int device, void* stream) {
  if (!host_dst || !dev_src || bytes < 0) return fail(B200W_E_INVALID, "dev_download: bad args");
  if (int e = ensure_device(device)) return e;
  HostPin pin;
  pin.pin(host_dst, (size_t)bytes);  // no-op for a buffer of the pinned pool
  B200W_CUDA(cudaMemcpyAsync(host_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  B200W_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}
extern "C" int b200w_dev_upload(void* dev_dst, const void* host_src, int64_t bytes, int device, void* stream) {
  if (!dev_dst || !host_src || bytes < 0) return fail(B200W_E_INVALID, "dev_upload: bad args");
  if (int e = ensure_device(device)) return e;
  HostPin pin;
  pin.pin(host_src, (size_t)bytes);
  B200W_CUDA(cudaMemcpyAsync(dev_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  B200W_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // the caller's pages are unpinned on return
  return 0;
}

extern "C" void* b200w_host_alloc(int64_t bytes) { return bytes < 0 ? nullptr : g_pool.alloc((size_t)bytes); }
extern "C" void b200w_host_free(void* p) { g_pool.free(p); }
extern "C" void b200w_host_trim(void) { g_pool.trim(); }
extern "C" void b200w_dev_trim(void) { g_dev_cache.trim(); }

int b200w_dev_intramodular(const double* dA, int64_t n, const int32_t* d_labels,
                           const uint8_t* d_small, double* d_total, double* d_within,
                           cudaStream_t st);

extern "C" int b200w_version(void) { return B200W_VERSION; }
extern "C" const char* b200w_last_error(void) { return t_error.c_str(); }
extern "C" int64_t b200w_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" double b200w_last_kernel_ms(void) { return g_last_kernel_ms; }

extern "C" int b200w_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < count; ++d) {
    int mj = 0;
    if (cudaDeviceGetAttribute(&mj, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && mj == 10)
      ++ok;
  }
  return ok;
}

extern "C" int b200w_connectivity_sweep(const double* X, int64_t m, int64_t n, const double* steps,
                                        int P, int net_type, double* k_out, uint8_t* bad_out,
                                        int device) {
  if (!X || !steps || !k_out || m < 1 || n < 1 || P < 1)
    return fail(B200W_E_INVALID, "connectivity_sweep: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  DevBuf dX, dZ, dBad, dK;
  StreamDrain drain{st.s};
  const int64_t ldk = b200w_padded_k(m);
  B200W_CUDA(dX.alloc((size_t)m * n * 8));
  B200W_CUDA(dZ.alloc((size_t)n * ldk * 8));
  B200W_CUDA(dBad.alloc((size_t)n));
  B200W_CUDA(dK.alloc((size_t)P * n * 8));
  B200W_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)m * n * 8, cudaMemcpyHostToDevice, st.s));
  if (int e = b200w_dev_standardize(dX.as<double>(), m, n, dZ.as<double>(), dBad.as<uint8_t>(),
                                    device, st.s))
    return e;
  if (int e = b200w_dev_sweep(dZ.as<double>(), dBad.as<uint8_t>(), m, n, steps, P, net_type, 0, n,
                              dK.as<double>(), device, st.s))
    return e;
  B200W_CUDA(cudaMemcpyAsync(k_out, dK.p, (size_t)P * n * 8, cudaMemcpyDeviceToHost, st.s));
  if (bad_out) B200W_CUDA(cudaMemcpyAsync(bad_out, dBad.p, (size_t)n, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

static int soft_threshold_stats_impl(const double* X, int64_t m, int64_t n, const double* steps, int P, int net_type,
                                     int nbreaks, double* stats_out, int64_t* buckets_out, int device, int gene_major);

extern "C" int b200w_soft_threshold_stats(const double* X, int64_t m, int64_t n, const double* steps,
                                          int P, int net_type, int nbreaks, double* stats_out,
                                          int64_t* buckets_out, int device) {
  return soft_threshold_stats_impl(X, m, n, steps, P, net_type, nbreaks, stats_out, buckets_out, device, 0);
}
extern "C" int b200w_soft_threshold_stats_gm(const double* Xt, int64_t m, int64_t n, const double* steps,
                                             int P, int net_type, int nbreaks, double* stats_out,
                                             int64_t* buckets_out, int device) {
  return soft_threshold_stats_impl(Xt, m, n, steps, P, net_type, nbreaks, stats_out, buckets_out, device, 1);
}

static int standardize_any(const double* dX, int64_t m, int64_t n, double* dZ, uint8_t* dBad, int device, void* s,
                           int gene_major) {
  return gene_major ? b200w_dev_standardize_gm(dX, m, n, dZ, dBad, device, s)
                    : b200w_dev_standardize(dX, m, n, dZ, dBad, device, s);
}

static int soft_threshold_stats_impl(const double* X, int64_t m, int64_t n, const double* steps, int P, int net_type,
                                     int nbreaks, double* stats_out, int64_t* buckets_out, int device,
                                     int gene_major) {
  if (!X || !steps || !stats_out || !buckets_out || m < 1 || n < 1 || P < 1 || nbreaks < 1 || nbreaks > 16)
    return fail(B200W_E_INVALID, "soft_threshold_stats: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  DevBuf dX, dZ, dBad, dK, dS, dB;
  StreamDrain drain{st.s};
  const int64_t ldk = b200w_padded_k(m);
  const size_t sbytes = (size_t)P * (4 + 2 * nbreaks) * 8;
  const size_t bbytes = (size_t)P * b200w_sft_exp_buckets() * 2 * 8;
  B200W_CUDA(dX.alloc((size_t)m * n * 8));
  B200W_CUDA(dZ.alloc((size_t)n * ldk * 8));
  B200W_CUDA(dBad.alloc((size_t)n));
  B200W_CUDA(dK.alloc((size_t)P * n * 8));
  B200W_CUDA(dS.alloc(sbytes));
  B200W_CUDA(dB.alloc(bbytes));
  B200W_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)m * n * 8, cudaMemcpyHostToDevice, st.s));
  if (int e = standardize_any(dX.as<double>(), m, n, dZ.as<double>(), dBad.as<uint8_t>(), device, st.s, gene_major))
    return e;
  if (int e = b200w_dev_sweep(dZ.as<double>(), dBad.as<uint8_t>(), m, n, steps, P, net_type, 0, n,
                              dK.as<double>(), device, st.s))
    return e;
  if (int e = b200w_dev_sft_stats(dK.as<double>(), P, n, n, nbreaks, dS.as<double>(), dB.as<int64_t>(),
                                  device, st.s))
    return e;
  B200W_CUDA(cudaMemcpyAsync(stats_out, dS.p, sbytes, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaMemcpyAsync(buckets_out, dB.p, bbytes, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

static int adjacency_host_impl(const double* X, int64_t m, int64_t n, int net_type, double power, int64_t row0,
                               int64_t nrows, double* A_out, int device, int gene_major);

extern "C" int b200w_adjacency(const double* X, int64_t m, int64_t n, int net_type, double power,
                               int64_t row0, int64_t nrows, double* A_out, int device) {
  return adjacency_host_impl(X, m, n, net_type, power, row0, nrows, A_out, device, 0);
}
extern "C" int b200w_adjacency_gm(const double* Xt, int64_t m, int64_t n, int net_type, double power,
                                  int64_t row0, int64_t nrows, double* A_out, int device) {
  return adjacency_host_impl(Xt, m, n, net_type, power, row0, nrows, A_out, device, 1);
}

static int adjacency_host_impl(const double* X, int64_t m, int64_t n, int net_type, double power, int64_t row0,
                               int64_t nrows, double* A_out, int device, int gene_major) {
  if (!X || !A_out || m < 1 || n < 1 || row0 < 0 || nrows < 1 || row0 + nrows > n)
    return fail(B200W_E_INVALID, "adjacency: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  Trace tr("b200w_adjacency", st.s);
  DevBuf dX, dZ, dBad, dA;
  HostPin pin_out;
  StreamDrain drain{st.s};
  pin_out.pin(A_out, (size_t)nrows * n * 8);
  tr.mark("pin");
  const int64_t ldk = b200w_padded_k(m);
  B200W_CUDA(dX.alloc((size_t)m * n * 8));
  B200W_CUDA(dZ.alloc((size_t)n * ldk * 8));
  B200W_CUDA(dBad.alloc((size_t)n));
  B200W_CUDA(dA.alloc((size_t)nrows * n * 8));
  tr.mark("alloc");
  B200W_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)m * n * 8, cudaMemcpyHostToDevice, st.s));
  tr.mark("h2d");
  if (int e = standardize_any(dX.as<double>(), m, n, dZ.as<double>(), dBad.as<uint8_t>(), device, st.s, gene_major))
    return e;
  if (int e = b200w_dev_adjacency(dZ.as<double>(), dBad.as<uint8_t>(), m, n, net_type, power, row0,
                                  nrows, dA.as<double>(), device, st.s))
    return e;
  tr.mark("kernels");
  B200W_CUDA(cudaMemcpyAsync(A_out, dA.p, (size_t)nrows * n * 8, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  tr.mark("d2h");
  return 0;
}

extern "C" int b200w_check_adjacency(const double* A, int64_t n, double* stats, int device) {
  if (!A || !stats || n < 1) return fail(B200W_E_INVALID, "check_adjacency: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  DevBuf dA, dS;
  StreamDrain drain{st.s};
  B200W_CUDA(dA.alloc((size_t)n * n * 8));
  B200W_CUDA(dS.alloc(6 * 8));
  B200W_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)n * n * 8, cudaMemcpyHostToDevice, st.s));
  if (int e = b200w_dev_check_adjacency(dA.as<double>(), n, dS.as<double>(), device, st.s)) return e;
  B200W_CUDA(cudaMemcpyAsync(stats, dS.p, 6 * 8, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

// Whole-matrix int8 contraction whose result goes to the host while the kernel is still running: the kernel
// reports finished row blocks through host-mapped flags (symmetric schedule, row-major over super-blocks) and
// this thread copies each block out on a second stream (at C3 20 GB of D2H take 0.39 s, the kernel 0.49 s).
static int tom_compute_to_host(const double* dA, const void* dOp, int64_t op_rows, const double* dScale,
                               const double* dK, int64_t n, int tom_type, int tom_denom, int out_mode, int w_min,
                               const TomColumnOperand* colp, double* dOut, double* out_host, cudaStream_t s) {
  TomProgress prog;
  if (int e = b200w_tom_compute_i8(dA, dOp, op_rows, dScale, dK, n, 0, n, tom_type, tom_denom, out_mode, dOut, s, 1, 0,
                                   0, nullptr, &prog, w_min, colp))
    return e;
  if (prog.nb > 0) {
    Stream s2;
    B200W_CUDA(s2.create());
    for (int b = 0; b < prog.nb; ++b) {
      unsigned long long spins = 0;
      while (prog.flags[b] == 0u) {
        if ((++spins & 0xfff) == 0 && cudaStreamQuery(s) != cudaErrorNotReady) break;  // kernel over (or failed)
      }
      const int64_t r0 = (int64_t)b * prog.rows_per_block;
      const int64_t r1 = r0 + prog.rows_per_block < n ? r0 + prog.rows_per_block : n;
      if (r1 > r0)
        B200W_CUDA(cudaMemcpyAsync(out_host + r0 * n, dOut + r0 * n, (size_t)(r1 - r0) * n * 8,
                                   cudaMemcpyDeviceToHost, s2.s));
    }
    B200W_CUDA(cudaStreamSynchronize(s));  // surfaces a kernel failure; all flags were set if it succeeded
    B200W_CUDA(cudaStreamSynchronize(s2.s));
    return 0;
  }
  B200W_CUDA(cudaMemcpyAsync(out_host, dOut, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s));
  B200W_CUDA(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int b200w_dev_tom_compute_to_host(const double* dA, const void* d_operand, int64_t op_rows,
                                             const double* d_scale, const double* d_k, int64_t n, int tom_type,
                                             int tom_denom, int out_mode, int precision, double* d_out,
                                             double* out_host, int device, void* stream) {
  if (!dA || !d_operand || !d_scale || !d_k || !d_out || !out_host || n < 1)
    return fail(B200W_E_INVALID, "tom_compute_to_host: bad args");
  if (tom_type < 0 || tom_type > 1 || tom_denom < 0 || tom_denom > 1 || out_mode < 0 || out_mode > 2)
    return fail(B200W_E_INVALID, "tom_compute_to_host: bad enum");
  if (op_rows < round_up(n, 128) || op_rows % 128)
    return fail(B200W_E_INVALID, "tom_compute_to_host: op_rows must be a multiple of 128 >= n");
  precision = resolve_precision(precision);
  if (!prec_is_i8(precision)) return fail(B200W_E_UNSUPPORTED, "tom_compute_to_host: int8 precision only");
  if (int e = ensure_device(device)) return e;
  HostPin pin;
  pin.pin(out_host, (size_t)n * n * 8);
  StreamDrain drain{(cudaStream_t)stream};
  return tom_compute_to_host(dA, d_operand, op_rows, d_scale, d_k, n, tom_type, tom_denom, out_mode,
                             prec_w_min(precision), nullptr, d_out, out_host, (cudaStream_t)stream);
}

// shared tail of b200w_tom / b200w_network: dA holds the full n x n adjacency on the device
static int tom_from_device_adjacency(const double* dA, int64_t n, int tom_type, int tom_denom,
                                     int out_mode, int precision, int64_t row0, int64_t nrows,
                                     double* out_host, int device, cudaStream_t s, bool asymmetric = false,
                                     int nan_mode = 0) {
  DevBuf dOp, dScale, dK, dOut, dAt, dOpT, dScaleT, dKT;
  StreamDrain drain{s};
  precision = resolve_precision(precision);
  B200W_CUDA(dOp.alloc((size_t)b200w_tom_operand_bytes(n, precision)));
  B200W_CUDA(dScale.alloc((size_t)round_up(n, 128) * 8));
  B200W_CUDA(dK.alloc((size_t)round_up(n, 128) * 8));
  B200W_CUDA(cudaMemsetAsync(dOp.p, 0, (size_t)b200w_tom_operand_bytes(n, precision), s));
  const bool condensed = out_mode == B200W_OUT_DISS_COND || out_mode == B200W_OUT_DISS_R8_COND;
  if (condensed && (row0 != 0 || nrows != n))
    return fail(B200W_E_INVALID, "tom: condensed output needs the whole matrix");
  const size_t out_bytes = condensed ? (size_t)n * (n - 1) / 2 * 8 : (size_t)nrows * n * 8;
  B200W_CUDA(dOut.alloc(out_bytes));
  const int64_t op_rows = round_up(n, 128);
  if (int e = tom_prepare_impl(dA, n, 0, n, precision, dOp.p, op_rows, dScale.as<double>(), dK.as<double>(),
                               device, s, nan_mode))
    return e;
  // The kernels contract rows with rows and mirror tiles, which is A @ A only for a symmetric A.  An
  // adjacency that is not (TomSimilarityFromAdj takes anything; a NaN waves any matrix through
  // checkAdjMat) gets its column operand from A^T: L = A @ A, k_j = column sums (wgcna.py:1145-1147).
  TomColumnOperand col;
  col.nan_mode = nan_mode;
  if (asymmetric) {
    B200W_CUDA(dAt.alloc((size_t)n * n * 8));
    B200W_CUDA(dOpT.alloc((size_t)b200w_tom_operand_bytes(n, precision)));
    B200W_CUDA(dScaleT.alloc((size_t)round_up(n, 128) * 8));
    B200W_CUDA(dKT.alloc((size_t)round_up(n, 128) * 8));
    B200W_CUDA(cudaMemsetAsync(dOpT.p, 0, (size_t)b200w_tom_operand_bytes(n, precision), s));
    if (int e = tom_transpose(dA, n, dAt.as<double>(), s)) return e;
    if (int e = tom_prepare_impl(dAt.as<double>(), n, 0, n, precision, dOpT.p, op_rows, dScaleT.as<double>(),
                                 dKT.as<double>(), device, s, nan_mode))
      return e;
    col.operand = dOpT.p;
    col.scale = dScaleT.as<double>();
    col.k = dKT.as<double>();
  } else {
    col.operand = dOp.p;
    col.scale = dScale.as<double>();
    col.k = dK.as<double>();
  }
  const TomColumnOperand* colp = (asymmetric || nan_mode) ? &col : nullptr;
  if (prec_is_i8(precision) && row0 == 0 && nrows == n && !condensed && !asymmetric &&
      !getenv("B200W_NO_OVERLAP")) {
    // whole matrix on the tensor-core path: the kernel reports finished row blocks through host-mapped
    // flags and they are copied out on a second stream while later blocks are still being computed
    // (3.2 GB of D2H at C2 take 62 ms, the kernel 35 ms: the copy hides the kernel, not the reverse)
    if (tom_type < 0 || tom_type > 1 || tom_denom < 0 || tom_denom > 1 || out_mode < 0 || out_mode > 2)
      return fail(B200W_E_INVALID, "tom: bad enum");
    return tom_compute_to_host(dA, dOp.p, op_rows, dScale.as<double>(), dK.as<double>(), n, tom_type, tom_denom,
                               out_mode, prec_w_min(precision), colp, dOut.as<double>(), out_host, s);
  }
  if (int e = tom_compute_impl(dA + row0 * n, dOp.p, op_rows, dScale.as<double>(), dK.as<double>(), n, row0,
                               nrows, tom_type, tom_denom, out_mode, precision, dOut.as<double>(), device, s, colp))
    return e;
  B200W_CUDA(cudaMemcpyAsync(out_host, dOut.p, out_bytes, cudaMemcpyDeviceToHost, s));
  B200W_CUDA(cudaStreamSynchronize(s));
  return 0;
}

extern "C" int b200w_tom(const double* A, int64_t n, int tom_type, int tom_denom, int out_mode,
                         int precision, int64_t row0, int64_t nrows, double* out, int check,
                         double* stats4, int device) {
  if (!A || !out || n < 1 || row0 < 0 || nrows < 1 || row0 + nrows > n)
    return fail(B200W_E_INVALID, "tom: bad args");
  if (tom_type < 0 || tom_type > 1) return fail(B200W_E_INVALID, "tom: bad TOM type");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  Trace tr("b200w_tom", st.s);
  DevBuf dA;
  HostPin pin_in, pin_out;
  StreamDrain drain{st.s};
  pin_in.pin(A, (size_t)n * n * 8);
  pin_out.pin(out, (out_mode == B200W_OUT_DISS_COND || out_mode == B200W_OUT_DISS_R8_COND)
                       ? (size_t)n * (n - 1) / 2 * 8 : (size_t)nrows * n * 8);
  tr.mark("pin");
  B200W_CUDA(dA.alloc((size_t)n * n * 8));
  tr.mark("alloc");
  B200W_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)n * n * 8, cudaMemcpyHostToDevice, st.s));
  tr.mark("h2d");
  // the checkAdjMat reductions run in every case: with check they decide (wgcna.py:1135-1138), without
  // they only tell whether the matrix is symmetric enough for the mirrored schedule
  double hs[6];
  {
    DevBuf dS;
    B200W_CUDA(dS.alloc(6 * 8));
    if (int e = b200w_dev_check_adjacency(dA.as<double>(), n, dS.as<double>(), device, st.s)) return e;
    B200W_CUDA(cudaMemcpyAsync(hs, dS.p, 6 * 8, cudaMemcpyDeviceToHost, st.s));
    B200W_CUDA(cudaStreamSynchronize(st.s));
  }
  if (stats4) for (int i = 0; i < 4; ++i) stats4[i] = hs[i];
  if (check && hs[3] == 0.0) {  // any NaN makes np.max / np.min NaN and both comparisons False
    if (hs[0] > 1e-12) return B200W_CHECK_ASYMMETRIC;
    const double lo = (tom_type == B200W_TOM_SIGNED) ? -1.0 : 0.0;
    if (hs[1] < lo || hs[2] > 1.0) return B200W_CHECK_RANGE;
  }
  tr.mark("check");
  // check (TOMsimilarity): NaN -> 0 first (:1185), then TomSimilarityFromAdj on whatever that leaves.
  // no check (TomSimilarityFromAdj called directly): NaNs stay and propagate like numpy's A @ A and sums.
  const int nan_mode = (!check && hs[3] > 0.0) ? 1 : 0;
  const bool asymmetric = hs[4] > 1e-12 || (nan_mode && hs[5] > 0.0);
  const int rc = tom_from_device_adjacency(dA.as<double>(), n, tom_type, tom_denom, out_mode, precision,
                                           row0, nrows, out, device, st.s, asymmetric, nan_mode);
  tr.mark("prepare+contraction+d2h");
  return rc;
}

extern "C" int b200w_network(const double* X, int64_t m, int64_t n, int net_type, double power,
                             int tom_type, int tom_denom, int out_mode, int precision,
                             double* A_out, double* tom_out, int device) {
  if (!X || !tom_out || m < 1 || n < 1) return fail(B200W_E_INVALID, "network: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  DevBuf dX, dZ, dBad, dA;
  HostPin pin_a, pin_t;
  StreamDrain drain{st.s};
  pin_a.pin(A_out, (size_t)n * n * 8);
  pin_t.pin(tom_out, (size_t)n * n * 8);
  const int64_t ldk = b200w_padded_k(m);
  B200W_CUDA(dX.alloc((size_t)m * n * 8));
  B200W_CUDA(dZ.alloc((size_t)n * ldk * 8));
  B200W_CUDA(dBad.alloc((size_t)n));
  B200W_CUDA(dA.alloc((size_t)n * n * 8));
  B200W_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)m * n * 8, cudaMemcpyHostToDevice, st.s));
  if (int e = b200w_dev_standardize(dX.as<double>(), m, n, dZ.as<double>(), dBad.as<uint8_t>(),
                                    device, st.s))
    return e;
  if (int e = b200w_dev_adjacency(dZ.as<double>(), dBad.as<uint8_t>(), m, n, net_type, power, 0, n,
                                  dA.as<double>(), device, st.s))
    return e;
  if (A_out)
    B200W_CUDA(cudaMemcpyAsync(A_out, dA.p, (size_t)n * n * 8, cudaMemcpyDeviceToHost, st.s));
  return tom_from_device_adjacency(dA.as<double>(), n, tom_type, tom_denom, out_mode, precision, 0,
                                   n, tom_out, device, st.s);
}

extern "C" int b200w_cross_correlation(const double* X, int64_t m, int64_t n, const double* Y, int64_t M,
                                       double* C_out, int device) {
  if (!X || !Y || !C_out || m < 1 || n < 1 || M < 1) return fail(B200W_E_INVALID, "cross_correlation: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  DevBuf dX, dY, dZx, dZy, dBx, dBy, dC;
  StreamDrain drain{st.s};
  const int64_t ldk = b200w_padded_k(m);
  B200W_CUDA(dX.alloc((size_t)m * n * 8));
  B200W_CUDA(dY.alloc((size_t)m * M * 8));
  B200W_CUDA(dZx.alloc((size_t)n * ldk * 8));
  B200W_CUDA(dZy.alloc((size_t)M * ldk * 8));
  B200W_CUDA(dBx.alloc((size_t)n));
  B200W_CUDA(dBy.alloc((size_t)M));
  B200W_CUDA(dC.alloc((size_t)n * M * 8));
  B200W_CUDA(cudaMemcpyAsync(dX.p, X, (size_t)m * n * 8, cudaMemcpyHostToDevice, st.s));
  B200W_CUDA(cudaMemcpyAsync(dY.p, Y, (size_t)m * M * 8, cudaMemcpyHostToDevice, st.s));
  if (int e = b200w_dev_standardize(dX.as<double>(), m, n, dZx.as<double>(), dBx.as<uint8_t>(), device, st.s)) return e;
  if (int e = b200w_dev_standardize(dY.as<double>(), m, M, dZy.as<double>(), dBy.as<uint8_t>(), device, st.s)) return e;
  if (int e = b200w_dev_cross_correlation(dZx.as<double>(), dBx.as<uint8_t>(), n, dZy.as<double>(), dBy.as<uint8_t>(),
                                          M, m, dC.as<double>(), device, st.s))
    return e;
  B200W_CUDA(cudaMemcpyAsync(C_out, dC.p, (size_t)n * M * 8, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}

extern "C" int b200w_linkage_average(const double* y_condensed, int64_t n, double* Z_out, int device) {
  if (!y_condensed || !Z_out || n < 2) return fail(B200W_E_INVALID, "linkage_average: bad args");
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  Trace tr("b200w_linkage_average", st.s);
  DevBuf dY, dD, dZ;
  HostPin pin;
  StreamDrain drain{st.s};
  const size_t cbytes = (size_t)n * (n - 1) / 2 * 8;
  pin.pin(y_condensed, cbytes);
  B200W_CUDA(dY.alloc(cbytes));
  B200W_CUDA(dD.alloc((size_t)n * n * 8));
  B200W_CUDA(dZ.alloc((size_t)(n - 1) * 4 * 8));
  B200W_CUDA(cudaMemcpyAsync(dY.p, y_condensed, cbytes, cudaMemcpyHostToDevice, st.s));
  tr.mark("h2d");
  if (int e = b200w_dev_condensed_to_square(dY.as<double>(), n, dD.as<double>(), device, st.s)) return e;
  if (int e = b200w_dev_linkage_average(dD.as<double>(), n, dZ.as<double>(), device, st.s)) return e;
  tr.mark("nn-chain");
  B200W_CUDA(cudaMemcpyAsync(Z_out, dZ.p, (size_t)(n - 1) * 4 * 8, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  const int rc = b200w_linkage_finish(Z_out, n);
  tr.mark("sort+label");
  return rc;
}

extern "C" int b200w_intramodular(const double* A, int64_t n, const int32_t* labels, int nmod,
                                  const uint8_t* small, double* k_total, double* k_within,
                                  int device) {
  if (!A || !labels || !small || !k_total || !k_within || n < 1 || nmod < 1)
    return fail(B200W_E_INVALID, "intramodular: bad args");
  for (int64_t i = 0; i < n; ++i)
    if (labels[i] < 0 || labels[i] >= nmod)
      return fail(B200W_E_INVALID, "intramodular: label %d out of range", (int)labels[i]);
  if (int e = ensure_device(device)) return e;
  Stream st;
  B200W_CUDA(st.create());
  DevBuf dA, dL, dS, dT, dW;
  StreamDrain drain{st.s};
  B200W_CUDA(dA.alloc((size_t)n * n * 8));
  B200W_CUDA(dL.alloc((size_t)n * 4));
  B200W_CUDA(dS.alloc((size_t)nmod));
  B200W_CUDA(dT.alloc((size_t)n * 8));
  B200W_CUDA(dW.alloc((size_t)n * 8));
  B200W_CUDA(cudaMemcpyAsync(dA.p, A, (size_t)n * n * 8, cudaMemcpyHostToDevice, st.s));
  B200W_CUDA(cudaMemcpyAsync(dL.p, labels, (size_t)n * 4, cudaMemcpyHostToDevice, st.s));
  B200W_CUDA(cudaMemcpyAsync(dS.p, small, (size_t)nmod, cudaMemcpyHostToDevice, st.s));
  if (int e = b200w_dev_intramodular(dA.as<double>(), n, dL.as<int32_t>(), dS.as<uint8_t>(),
                                     dT.as<double>(), dW.as<double>(), st.s))
    return e;
  B200W_CUDA(cudaMemcpyAsync(k_total, dT.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaMemcpyAsync(k_within, dW.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st.s));
  B200W_CUDA(cudaStreamSynchronize(st.s));
  return 0;
}
