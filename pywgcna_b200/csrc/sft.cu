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


// ---- exactly rounded mean from the mantissa sums (statistics.mean, wgcna.py:930) ------------------------
// total = sum_e (lo_e + hi_e 2^27) 2^(e - e0) as an unsigned multi-word integer (positive and negative parts
// kept apart), mean = total 2^(e0 - kExpBias - 53) / n rounded once (half to even): long division of
// total 2^128 by n, top 53 bits of the quotient, round bit, sticky = lower quotient bits | remainder.
constexpr int kBigWords = 44;  // 2816 bits: 2304 exponents + 27 + 54 + 128 shift + head room

struct BigU {
  unsigned long long w[kBigWords];
  __device__ void clear() { for (int i = 0; i < kBigWords; ++i) w[i] = 0ull; }
  __device__ void add_shifted(unsigned long long v, int shift) {
    const int wi = shift >> 6, b = shift & 63;
    unsigned long long lo = v << b, hi = b ? (v >> (64 - b)) : 0ull;
    unsigned long long c = 0;
    for (int i = wi; i < kBigWords && (lo | hi | c); ++i) {
      const unsigned long long add = lo, old = w[i];
      unsigned long long sum = old + add;
      unsigned long long c2 = sum < old;
      sum += c;
      c2 += sum < c;
      w[i] = sum;
      c = c2;
      lo = hi;
      hi = 0ull;
    }
  }
  __device__ int cmp(const BigU& o) const {
    for (int i = kBigWords - 1; i >= 0; --i)
      if (w[i] != o.w[i]) return w[i] < o.w[i] ? -1 : 1;
    return 0;
  }
  __device__ void sub(const BigU& o) {  // this -= o, this >= o
    unsigned long long borrow = 0;
    for (int i = 0; i < kBigWords; ++i) {
      const unsigned long long a = w[i], b = o.w[i];
      const unsigned long long d = a - b - borrow;
      borrow = (a < b) || (a == b && borrow) ? 1ull : 0ull;
      w[i] = d;
    }
  }
  __device__ int bit_length() const {
    for (int i = kBigWords - 1; i >= 0; --i)
      if (w[i]) return i * 64 + (64 - __clzll((long long)w[i]));
    return 0;
  }
  __device__ unsigned long long bits(int lo_bit, int count) const {  // count <= 64 bits starting at lo_bit >= 0
    const int wi = lo_bit >> 6, b = lo_bit & 63;
    unsigned long long v = w[wi] >> b;
    if (b && wi + 1 < kBigWords) v |= w[wi + 1] << (64 - b);
    return count == 64 ? v : (v & ((1ull << count) - 1ull));
  }
  __device__ bool any_below(int bit) const {  // any set bit at positions < bit
    const int wi = bit >> 6, b = bit & 63;
    for (int i = 0; i < wi; ++i)
      if (w[i]) return true;
    return b ? (w[wi] & ((1ull << b) - 1ull)) != 0ull : false;
  }
};

// pos / neg: lists of (bucket index, lo, hi) with lo, hi >= 0 after splitting by sign
__device__ double exact_mean_from_buckets(const int* idx, const long long* lo, const long long* hi, int cnt,
                                          long long n, BigU& P, BigU& N) {
  if (cnt == 0) return 0.0;
  int e0 = idx[0];
  for (int i = 1; i < cnt; ++i) e0 = idx[i] < e0 ? idx[i] : e0;
  P.clear();
  N.clear();
  for (int i = 0; i < cnt; ++i) {
    const int sh = idx[i] - e0 + 128;  // the 2^128 of the division is folded in here
    if (lo[i] >= 0) P.add_shifted((unsigned long long)lo[i], sh); else N.add_shifted((unsigned long long)(-lo[i]), sh);
    if (hi[i] >= 0) P.add_shifted((unsigned long long)hi[i], sh + 27); else N.add_shifted((unsigned long long)(-hi[i]), sh + 27);
  }
  const int c = P.cmp(N);
  if (c == 0) return 0.0;
  const bool negative = c < 0;
  BigU& T = negative ? N : P;
  T.sub(negative ? P : N);
  // long division by n (< 2^20) in 32-bit limbs, quotient in place
  unsigned long long rem = 0;
  for (int i = kBigWords - 1; i >= 0; --i) {
    const unsigned long long word = T.w[i];
    unsigned long long cur = (rem << 32) | (word >> 32);
    const unsigned long long qh = cur / (unsigned long long)n;
    rem = cur % (unsigned long long)n;
    cur = (rem << 32) | (word & 0xffffffffull);
    const unsigned long long ql = cur / (unsigned long long)n;
    rem = cur % (unsigned long long)n;
    T.w[i] = (qh << 32) | ql;
  }
  const int bl = T.bit_length();  // >= 128 - 20 + 1
  unsigned long long mant = T.bits(bl - 53, 53);
  const bool round_bit = T.bits(bl - 54, 1) != 0ull;
  const bool sticky = T.any_below(bl - 54) || rem != 0ull;
  if (round_bit && (sticky || (mant & 1ull))) ++mant;  // half to even; 2^53 is still exact in a double
  // value = quotient 2^(e0 - kExpBias - 53 - 128), quotient ~ mant 2^(bl - 53)
  const double v = ldexp((double)mant, bl - 53 + e0 - kExpBias - 53 - 128);
  return negative ? -v : v;
}

// R^2 of the ten-point regression log10(p(k)) ~ 1 + log10(k) of scaleFreeFitIndex (wgcna.py:1017-1039), the
// only regression output the power selection (:945-953) reads; same cleaning of the bins as the host code.
// Closed form with centred sums -- the statsmodels pinv solution to rounding.
__device__ double scale_free_rsq(const double* cnt, const double* tot, double mn, double mx, long long n, int nb) {
  double x[kSftMaxBreaks], y[kSftMaxBreaks];
  const double step = (mx - mn) / (double)nb;
  double xb = 0.0, yb = 0.0;
  for (int b = 0; b < nb; ++b) {
    // np.linspace arithmetic (multiply, then add: no fused multiply-add), last point = mx
    const double g0 = __dadd_rn(__dmul_rn((double)b, step), mn);
    const double g1 = (b + 1 == nb) ? mx : __dadd_rn(__dmul_rn((double)(b + 1), step), mn);
    const double mid = 0.5 * (g1 + g0);                       // :1025-1027
    double dk = cnt[b] > 0.0 ? tot[b] / cnt[b] : mid;         // :1019, :1029-1030
    if (dk == 0.0) dk = mid;                                  // :1031-1032
    x[b] = log10(dk);                                         // :1035
    y[b] = log10(cnt[b] / (double)n + 1e-09);                 // :1022, :1036
    xb += x[b];
    yb += y[b];
  }
  xb /= nb;
  yb /= nb;
  double sxx = 0.0, sxy = 0.0, syy = 0.0;
  for (int b = 0; b < nb; ++b) {
    const double dx = x[b] - xb, dy = y[b] - yb;
    sxx += dx * dx;
    sxy += dx * dy;
    syy += dy * dy;
  }
  const double slope = sxx > 0.0 ? sxy / sxx : 0.0;
  double ssr = 0.0;
  for (int b = 0; b < nb; ++b) {
    const double r = (y[b] - yb) - slope * (x[b] - xb);
    ssr += r * r;
  }
  return 1.0 - ssr / syy;  // NaN when the bins are degenerate, like the reference
}

// stats row layout: [0] min [1] max [2] median [3] plain float64 sum (informational) [4..4+nb) counts
// [4+nb..4+2nb) bin sums
// cached != 0: the row is first copied into dynamic shared memory and every pass reads it from there
__global__ void __launch_bounds__(kSftThreads)
sft_stats_kernel(const double* __restrict__ k, int64_t n, int64_t ldk, int nb, int cached,
                 double* __restrict__ stats, long long* __restrict__ buckets, double* __restrict__ fit) {
  extern __shared__ __align__(16) double dyn_smem[];
  // dynamic shared memory: [kExpBuckets][kPieces] int sums of the mantissa pieces, then the optional row cache
  int (*bk)[kPieces] = reinterpret_cast<int (*)[kPieces]>(dyn_smem);
  double* row_cache = dyn_smem + kExpBuckets * kPieces * sizeof(int) / sizeof(double);
  __shared__ double red_d[kSftThreads / 32];
  __shared__ long long red_l[kSftThreads / 32];
  __shared__ __align__(8) unsigned int hist[2048];
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
  if (fit == nullptr) return;  // block-uniform
  // what the power selection needs, still on the device: R^2 of the first regression and the exactly rounded
  // mean (the non-empty exponent buckets are compacted by all threads, thread 0 does the multi-word part)
  __shared__ int s_cnt;
  __shared__ int s_idx[192];
  __shared__ long long s_lo[192], s_hi[192];
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < kExpBuckets; i += kSftThreads) {
    const long long lo = (long long)bk[i][0] + (long long)bk[i][1] * (1LL << 11) + (long long)bk[i][2] * (1LL << 22);
    const long long hi = (long long)bk[i][3] * (1LL << 6) + (long long)bk[i][4] * (1LL << 17);
    if (lo != 0 || hi != 0) {
      const int at = atomicAdd(&s_cnt, 1);
      if (at < 192) { s_idx[at] = i; s_lo[at] = lo; s_hi[at] = hi; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double mean;
    if (!isfinite(plain)) mean = plain / (double)n;
    else if (s_cnt > 192) mean = plain / (double)n;  // > 192 distinct binades: not a connectivity vector
    else {
      BigU* big = reinterpret_cast<BigU*>(hist);  // 2 x 352 B inside the 8 KB radix histogram, idle by now
      mean = exact_mean_from_buckets(s_idx, s_lo, s_hi, s_cnt, (long long)n, big[0], big[1]);
    }
    fit[(int64_t)p * 2] = scale_free_rsq(out + 4, out + 4 + nb, mn, mx, (long long)n, nb);
    fit[(int64_t)p * 2 + 1] = mean;
  }
}

// Power selection of pickSoftThreshold (wgcna.py:945-953) from fit[p] = (R^2, mean k): the smallest power with
// R^2 > RsquaredCut and mean(k) <= MeanCut, else powers[argsort(R^2)[-1]] (raw R^2, not the slope-signed one;
// NaN sorts last in numpy, so a NaN R^2 wins that branch as it does in the reference).
__global__ void sft_select_kernel(const double* __restrict__ fit, const double* __restrict__ powers, int P,
                                  double rsq_cut, double mean_cut, double* __restrict__ power_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int pick = -1;
  for (int p = 0; p < P && pick < 0; ++p)
    if (fit[p * 2] > rsq_cut && fit[p * 2 + 1] <= mean_cut) pick = p;
  if (pick < 0) {
    pick = 0;
    for (int p = 1; p < P; ++p) {
      const double a = fit[p * 2], b = fit[pick * 2];
      if (isnan(a) || (!isnan(b) && a >= b)) pick = p;
    }
  }
  power_out[0] = powers[pick];
  power_out[1] = (double)pick;
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_sft_exp_buckets(void) { return kExpBuckets; }

static int sft_stats_impl(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks, double* d_stats,
                          int64_t* d_buckets, double* d_fit, int device, void* stream);

extern "C" int b200w_dev_sft_stats(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks, double* d_stats,
                                   int64_t* d_buckets, int device, void* stream) {
  return sft_stats_impl(d_k, P, n, ldk, nbreaks, d_stats, d_buckets, nullptr, device, stream);
}

extern "C" int b200w_dev_sft_select(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks,
                                    const double* d_powers, double rsq_cut, double mean_cut, double* d_stats,
                                    int64_t* d_buckets, double* d_fit, double* d_power_out, int device,
                                    void* stream) {
  if (!d_powers || !d_fit || !d_power_out) return fail(B200W_E_INVALID, "sft_select: bad args");
  if (int e = sft_stats_impl(d_k, P, n, ldk, nbreaks, d_stats, d_buckets, d_fit, device, stream)) return e;
  sft_select_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_fit, d_powers, P, rsq_cut, mean_cut, d_power_out);
  B200W_LAUNCHED();
  return 0;
}

static int sft_stats_impl(const double* d_k, int P, int64_t n, int64_t ldk, int nbreaks, double* d_stats,
                          int64_t* d_buckets, double* d_fit, int device, void* stream) {
  if (!d_k || !d_stats || !d_buckets || P < 1 || n < 1 || ldk < n || nbreaks < 1 || nbreaks > kSftMaxBreaks)
    return fail(B200W_E_INVALID, "sft_stats: bad args");
  if (int e = ensure_device(device)) return e;
  if (n >= (1LL << 20)) return fail(B200W_E_UNSUPPORTED, "sft_stats: more than 2^20 - 1 genes");
  // rows up to ~21k connectivities are staged in shared memory (one global read instead of ten)
  static const size_t kCacheMax = 164 * 1024, kBucketBytes = (size_t)kExpBuckets * kPieces * sizeof(int);
  static_assert(kBucketBytes % sizeof(double) == 0, "row cache alignment");
  const size_t row_bytes = (size_t)n * sizeof(double);
  const int cached = row_bytes <= kCacheMax ? 1 : 0;
  B200W_CUDA(cudaFuncSetAttribute(sft_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(kBucketBytes + kCacheMax)));
  sft_stats_kernel<<<P, kSftThreads, kBucketBytes + (cached ? row_bytes : 0), (cudaStream_t)stream>>>(
      d_k, n, ldk, nbreaks, cached, d_stats, (long long*)d_buckets, d_fit);
  B200W_LAUNCHED();
  return 0;
}
