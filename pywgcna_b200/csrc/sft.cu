// Per-power statistics of the connectivities for the scale-free topology fit, on the device:
// everything WGCNA.pickSoftThreshold derives from k[:, power] except the two 10-point regressions
// (PyWGCNA/wgcna.py:924-932 and scaleFreeFitIndex :1017-1034):
//   min, max                       (:1018 pd.cut range, :932)
//   median                         statistics.median (:931): exact k-th order statistics by radix select
//   histogram                      pd.cut(k, nBreaks) -> per-bin count and sum (:1018-1022), bins
//                                  (e_i, e_i+1] with numpy's linspace arithmetic and the lowered first edge
//   exact sum                      statistics.mean (:930) is the correctly rounded exact mean; every double
//                                  is M * 2^e with a 53-bit integer M: the kernel adds the five 11-bit pieces
//                                  of M into int32 counters per exponent (integer atomics: exact, order free)
//                                  and hands them over as (lo, hi) int64 pairs meaning lo + hi * 2^27; the
//                                  host combines the ~50 non-empty pairs as Python integers.
// One CTA per power; a row of k (160 KB at C2) is staged in shared memory once and read by all passes --
// a latency-sized kernel, not a bandwidth one: it removes the 3.2 MB D2H + ~8 ms of host sorting from
// every step.
#include "common.cuh"
#include <climits>

namespace b200w {

constexpr int kSftThreads = 256;
constexpr int kSftMaxBreaks = 16;
constexpr int kExpBuckets = 2304;  // frexp exponent + 1100 in [27, 2124]
constexpr int kExpBias = 1100;
constexpr int kPieces = 5;

__device__ __forceinline__ unsigned long long sortable_key(double x) {
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
  unsigned long long u = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

template <class T, class Op>
__device__ __forceinline__ T block_reduce(T v, T* red, Op op) {
  __syncthreads();
  for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  T r = red[0];
  for (int w = 1; w < kSftThreads / 32; ++w) r = op(r, red[w]);  // fixed order -> deterministic
  return r;
}

// hist[bin] += 1 with the lanes of a warp that hit the same bin combined first: the top radix digit (sign +
// exponent) and the exponent buckets put almost every element of a row into two or three bins
__device__ __forceinline__ void hist_add(unsigned int* hist, unsigned int bin, bool active) {
  const unsigned int live = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const unsigned int peers = __match_any_sync(live, bin);
  if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned int)__popc(peers));
}

// rank-r order statistic (0-based) of row[0..n) by 11-bit radix select on the sortable keys.  The bin that
// holds the rank is found by a block-wide prefix sum over the 2048 counters (8 per thread).
__device__ double radix_select(const double* __restrict__ row, int64_t n, int64_t r, unsigned int* hist) {
  unsigned long long prefix = 0, mask = 0;
  __shared__ unsigned int s_warp_tot[kSftThreads / 32];
  __shared__ unsigned int s_bin;
  __shared__ long long s_rank;
  constexpr int kPer = 2048 / kSftThreads;  // counters per thread
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int shift = 53; shift >= -2; shift -= 11) {
    const int sh = shift < 0 ? 0 : shift;
    const int bits = shift < 0 ? 11 + shift : 11;  // last pass covers the remaining low bits
    const unsigned int nb = 1u << bits;
    for (unsigned int i = threadIdx.x; i < 2048; i += kSftThreads) hist[i] = 0;
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += kSftThreads) {  // whole warps stay in the loop for the ballots
      const int64_t i = i0 + threadIdx.x;
      unsigned long long k = 0;
      bool hit = false;
      if (i < n) {
        k = sortable_key(row[i]);
        hit = (k & mask) == prefix;
      }
      hist_add(hist, (unsigned int)((k >> sh) & (nb - 1)), hit);
    }
    __syncthreads();
    // exclusive prefix over the counters: thread t owns bins [t * kPer, (t + 1) * kPer)
    unsigned int mine[kPer], tot = 0;
#pragma unroll
    for (int q = 0; q < kPer; ++q) { mine[q] = hist[threadIdx.x * kPer + q]; tot += mine[q]; }
    unsigned int incl = tot;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    unsigned int before = incl - tot;
    for (int w = 0; w < warp; ++w) before += s_warp_tot[w];
    // exactly one thread has before <= r < before + tot
    if ((long long)before <= r && r < (long long)before + (long long)tot) {
      long long rr = r - before;
      int q = 0;
      for (; q < kPer - 1; ++q) {
        if (rr < (long long)mine[q]) break;
        rr -= mine[q];
      }
      s_bin = threadIdx.x * kPer + q;
      s_rank = rr;
    }
    __syncthreads();
    prefix |= (unsigned long long)s_bin << sh;
    r = s_rank;
    mask |= (unsigned long long)(nb - 1) << sh;
    __syncthreads();
  }
  return key_to_double(prefix);
}

// stats row layout: [0] min [1] max [2] median [3] plain float64 sum (informational) [4..4+nb) counts
// [4+nb..4+2nb) bin sums
// cached != 0: the row is first copied into dynamic shared memory and every pass reads it from there
__global__ void __launch_bounds__(kSftThreads)
sft_stats_kernel(const double* __restrict__ k, int64_t n, int64_t ldk, int nb, int cached,
                 double* __restrict__ stats, long long* __restrict__ buckets) {
  extern __shared__ __align__(16) double dyn_smem[];
  // dynamic shared memory: [kExpBuckets][kPieces] int sums of the mantissa pieces, then the optional row cache
  int (*bk)[kPieces] = reinterpret_cast<int (*)[kPieces]>(dyn_smem);
  double* row_cache = dyn_smem + kExpBuckets * kPieces * sizeof(int) / sizeof(double);
  __shared__ double red_d[kSftThreads / 32];
  __shared__ long long red_l[kSftThreads / 32];
  __shared__ unsigned int hist[2048];
  __shared__ double edges[kSftMaxBreaks + 1];
  const int p = blockIdx.x;
  const double* row = k + (int64_t)p * ldk;
  double* out = stats + (int64_t)p * (4 + 2 * nb);

  double mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += kSftThreads) {
    const double v = row[i];
    if (cached) row_cache[i] = v;
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  if (cached) row = row_cache;  // (the reductions below start with a __syncthreads)
  mn = block_reduce(mn, red_d, [](double a, double b) { return fmin(a, b); });
  mx = block_reduce(mx, red_d, [](double a, double b) { return fmax(a, b); });

  // pd.cut(k, nb) edges (pandas 2.2 _nbins_to_bins): np.linspace(mn, mx, nb + 1) = arange * step + mn with
  // the last point set to mx, then edges[0] -= 0.001 * (mx - mn); for mn == mx the range is widened first.
  if (threadIdx.x == 0) {
    double lo = mn, hi = mx;
    if (mn == mx) {
      lo = mn - (mn != 0.0 ? 0.001 * fabs(mn) : 0.001);
      hi = mx + (mx != 0.0 ? 0.001 * fabs(mx) : 0.001);
    }
    const double step = __ddiv_rn(__dsub_rn(hi, lo), (double)nb);
    for (int i = 0; i < nb; ++i) edges[i] = __dadd_rn(__dmul_rn((double)i, step), lo);
    edges[nb] = hi;
    if (mn != mx) edges[0] = __dsub_rn(edges[0], __dmul_rn(__dsub_rn(mx, mn), 0.001));
  }
  for (int i = threadIdx.x; i < kExpBuckets * kPieces; i += kSftThreads) (&bk[0][0])[i] = 0;
  __syncthreads();

  // histogram (bin = #edges < k, minus 1: right-closed bins) + exact mantissa sums per exponent: the 53-bit
  // integer mantissa is cut into five 11-bit pieces that are added with native 32-bit shared-memory atomics
  // (11 bits x n < 2^31 for n < 2^20 genes; b200w_dev_sft_stats rejects more).  64-bit shared atomics are
  // compare-and-swap loops and took half of the kernel; warp-level pre-reduction by exponent
  // (match.any + redux) measured slower than letting the atomic unit resolve the conflicts.
  double cnt[kSftMaxBreaks], sum[kSftMaxBreaks];
#pragma unroll
  for (int b = 0; b < kSftMaxBreaks; ++b) { cnt[b] = 0.0; sum[b] = 0.0; }
  double plain = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += kSftThreads) {
    const double v = row[i];
    int bin = -1;
    for (int e = 0; e <= nb; ++e) bin += (edges[e] < v) ? 1 : 0;
    if (bin >= 0 && bin < nb) {
#pragma unroll
      for (int b = 0; b < kSftMaxBreaks; ++b)
        if (b == bin) { cnt[b] += 1.0; sum[b] += v; }
    }
    plain += v;
    if (v != 0.0 && isfinite(v)) {
      int ex;
      const double m = frexp(v, &ex);                            // v = m * 2^ex, 0.5 <= |m| < 1
      const long long M = (long long)(m * 9007199254740992.0);   // m * 2^53, exact
      const long long mag = M < 0 ? -M : M;
      const int sgn = M < 0 ? -1 : 1;
#pragma unroll
      for (int q = 0; q < kPieces; ++q)
        atomicAdd(&bk[ex + kExpBias][q], sgn * (int)((mag >> (11 * q)) & 0x7ff));
    }
  }
  auto add = [](double a, double b) { return a + b; };
  for (int b = 0; b < nb; ++b) {
    double c = 0.0, s = 0.0;
#pragma unroll
    for (int q = 0; q < kSftMaxBreaks; ++q)
      if (q == b) { c = cnt[q]; s = sum[q]; }
    c = block_reduce(c, red_d, add);
    s = block_reduce(s, red_d, add);
    if (threadIdx.x == 0) { out[4 + b] = c; out[4 + nb + b] = s; }
  }
  plain = block_reduce(plain, red_d, add);
  __syncthreads();
  for (int i = threadIdx.x; i < kExpBuckets; i += kSftThreads) {
    // the pair (lo, hi) the host reads stands for lo + hi * 2^27
    const long long lo = (long long)bk[i][0] + (long long)bk[i][1] * (1LL << 11) + (long long)bk[i][2] * (1LL << 22);
    const long long hi = (long long)bk[i][3] * (1LL << 6) + (long long)bk[i][4] * (1LL << 17);
    buckets[((int64_t)p * kExpBuckets + i) * 2] = lo;
    buckets[((int64_t)p * kExpBuckets + i) * 2 + 1] = hi;
  }
  // statistics.median: middle element, or the mean of the two middle elements.  The lower middle element
  // (rank n/2 - 1) is the upper one again when enough copies of it or smaller values sit below rank n/2,
  // otherwise the largest value below it: one counting pass instead of a second select.
  const double m_hi = radix_select(row, n, n / 2, hist);
  double med = m_hi;
  if ((n & 1) == 0) {
    long long less = 0;
    double below = -INFINITY;
    for (int64_t i = threadIdx.x; i < n; i += kSftThreads) {
      const double v = row[i];
      if (v < m_hi) { ++less; below = fmax(below, v); }
    }
    less = block_reduce(less, red_l, [](long long a, long long b) { return a + b; });
    below = block_reduce(below, red_d, [](double a, double b) { return fmax(a, b); });
    const double m_lo = (n / 2 - 1 >= less) ? m_hi : below;
    med = (m_lo + m_hi) / 2.0;
  }
  if (threadIdx.x == 0) { out[0] = mn; out[1] = mx; out[2] = med; out[3] = plain; }
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_sft_exp_buckets(void) { return kExpBuckets; }

extern "C" int b200w_dev_sft_stats(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks, double* d_stats,
                                   int64_t* d_buckets, int device, void* stream) {
  if (!d_k || !d_stats || !d_buckets || P < 1 || n < 1 || ldk < n || nbreaks < 1 || nbreaks > kSftMaxBreaks)
    return fail(B200W_E_INVALID, "sft_stats: bad args");
  if (int e = ensure_device(device)) return e;
  if (n >= (1LL << 20)) return fail(B200W_E_UNSUPPORTED, "sft_stats: more than 2^20 - 1 genes");
  // rows up to ~21k connectivities are staged in shared memory (one global read instead of ten)
  static const size_t kCacheMax = 170 * 1024, kBucketBytes = (size_t)kExpBuckets * kPieces * sizeof(int);
  static_assert(kBucketBytes % sizeof(double) == 0, "row cache alignment");
  const size_t row_bytes = (size_t)n * sizeof(double);
  const int cached = row_bytes <= kCacheMax ? 1 : 0;
  B200W_CUDA(cudaFuncSetAttribute(sft_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kBucketBytes + kCacheMax)));
  sft_stats_kernel<<<P, kSftThreads, kBucketBytes + (cached ? row_bytes : 0), (cudaStream_t)stream>>>(
      d_k, n, ldk, nbreaks, cached, d_stats, (long long*)d_buckets);
  B200W_LAUNCHED();
  return 0;
}
