// Per-power statistics of the connectivities for the scale-free topology fit, on the device:
// everything WGCNA.pickSoftThreshold derives from k[:, power] except the two 10-point regressions
// (PyWGCNA/wgcna.py:924-932 and scaleFreeFitIndex :1017-1034):
//   min, max                       (:1018 pd.cut range, :932)
//   median                         statistics.median (:931): exact k-th order statistics by radix select
//   histogram                      pd.cut(k, nBreaks) -> per-bin count and sum (:1018-1022), bins
//                                  (e_i, e_i+1] with numpy's linspace arithmetic and the lowered first edge
//   exact sum                      statistics.mean (:930) is the correctly rounded exact mean; every double
//                                  is M * 2^e with a 53-bit integer M: the kernel adds the 27-bit halves of
//                                  M into int64 counters per exponent (integer atomics: exact, order free),
//                                  the host combines the ~50 non-empty counters as Python integers.
// One CTA per power; k[P, n] is 3.2 MB at C2, so this is a latency-sized kernel (~30 us), not a
// bandwidth one -- it removes the 3.2 MB D2H + ~8 ms of host sorting from every step.
#include "common.cuh"

namespace b200w {

constexpr int kSftThreads = 256;
constexpr int kSftMaxBreaks = 16;
constexpr int kExpBuckets = 2304;  // frexp exponent + 1100 in [27, 2124]
constexpr int kExpBias = 1100;

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

// rank-r order statistic (0-based) of row[0..n) by 11-bit radix select on the sortable keys
__device__ double radix_select(const double* __restrict__ row, int64_t n, int64_t r, unsigned int* hist) {
  unsigned long long prefix = 0, mask = 0;
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_rank;
  for (int shift = 53; shift >= -2; shift -= 11) {
    const int sh = shift < 0 ? 0 : shift;
    const int bits = shift < 0 ? 11 + shift : 11;  // last pass covers the remaining low bits
    const unsigned int nb = 1u << bits;
    for (unsigned int i = threadIdx.x; i < nb; i += kSftThreads) hist[i] = 0;
    __syncthreads();
    for (int64_t i = threadIdx.x; i < n; i += kSftThreads) {
      const unsigned long long k = sortable_key(row[i]);
      if ((k & mask) == prefix) atomicAdd(&hist[(unsigned int)((k >> sh) & (nb - 1))], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long rr = r;
      unsigned int b = 0;
      for (; b < nb; ++b) {
        if (rr < (long long)hist[b]) break;
        rr -= hist[b];
      }
      s_prefix = prefix | ((unsigned long long)b << sh);
      s_rank = rr;
    }
    __syncthreads();
    prefix = s_prefix;
    r = s_rank;
    mask |= (unsigned long long)(nb - 1) << sh;
    __syncthreads();
  }
  return key_to_double(prefix);
}

// stats row layout: [0] min [1] max [2] median [3] plain float64 sum (informational) [4..4+nb) counts
// [4+nb..4+2nb) bin sums
__global__ void __launch_bounds__(kSftThreads)
sft_stats_kernel(const double* __restrict__ k, int64_t n, int64_t ldk, int nb, double* __restrict__ stats,
                 long long* __restrict__ buckets) {
  __shared__ double red_d[kSftThreads / 32];
  __shared__ unsigned int hist[2048];
  __shared__ long long bk_lo[kExpBuckets], bk_hi[kExpBuckets];
  __shared__ double edges[kSftMaxBreaks + 1];
  const int p = blockIdx.x;
  const double* row = k + (int64_t)p * ldk;
  double* out = stats + (int64_t)p * (4 + 2 * nb);

  double mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += kSftThreads) {
    const double v = row[i];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
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
  for (int i = threadIdx.x; i < kExpBuckets; i += kSftThreads) { bk_lo[i] = 0; bk_hi[i] = 0; }
  __syncthreads();

  // histogram (bin = #edges < k, minus 1: right-closed bins) + exact mantissa sums per exponent
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
      const double m = frexp(v, &ex);                       // v = m * 2^ex, 0.5 <= |m| < 1
      const long long M = (long long)(m * 9007199254740992.0);  // m * 2^53, exact
      const long long hi = M >> 27, lo = M - (hi << 27);
      atomicAdd((unsigned long long*)&bk_lo[ex + kExpBias], (unsigned long long)lo);
      atomicAdd((unsigned long long*)&bk_hi[ex + kExpBias], (unsigned long long)hi);
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
    buckets[((int64_t)p * kExpBuckets + i) * 2] = bk_lo[i];
    buckets[((int64_t)p * kExpBuckets + i) * 2 + 1] = bk_hi[i];
  }
  // statistics.median: middle element, or the mean of the two middle elements
  const double m_hi = radix_select(row, n, n / 2, hist);
  double med = m_hi;
  if ((n & 1) == 0) {
    const double m_lo = radix_select(row, n, n / 2 - 1, hist);
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
  sft_stats_kernel<<<P, kSftThreads, 0, (cudaStream_t)stream>>>(d_k, n, ldk, nbreaks, d_stats,
                                                                  (long long*)d_buckets);
  B200W_LAUNCHED();
  return 0;
}
