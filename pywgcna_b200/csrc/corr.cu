// Correlation side of the hot path: standardise, soft-threshold sweep, adjacency.
// Reference arithmetic: np.corrcoef (numpy lib/_function_base_impl.py cov/corrcoef) as called at
// PyWGCNA/wgcna.py:882 and :1097; sweep wgcna.py:884-916; adjacency wgcna.py:1102-1111.
#include "gemm_f64.cuh"

namespace b200w {

// -------------------------------------------------------------------------------------------
// K0  standardise.  X is m x n (samples x genes, row-major).  One thread per gene, a warp covers
// 32 consecutive genes so every sample row is read as one 256-byte coalesced segment.  The mean is
// the left-to-right sum over samples divided by m -- the order numpy's add.reduce uses along the
// strided axis of X.T -- so that exactly-constant genes centre to exact zeros as they do in
// np.cov (zero variance -> NaN row/column).  Output Z is gene-major (n x ldk),
// written through a 32 x 33 shared tile so the stores are coalesced as well.
// -------------------------------------------------------------------------------------------
constexpr int kStdWarps = 4;

__global__ void __launch_bounds__(kStdWarps * 32)
standardize_kernel(const double* __restrict__ X, int64_t m, int64_t n, int64_t ldk,
                   double* __restrict__ Z, uint8_t* __restrict__ bad) {
  __shared__ double tile[kStdWarps][32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t g0 = ((int64_t)blockIdx.x * kStdWarps + w) * 32;
  if (g0 >= n) return;  // warp-uniform
  const int64_t g = g0 + lane;
  const bool live = g < n;
  const double* col = X + (live ? g : n - 1);

  double sum = 0.0;
  bool finite = true;
  for (int64_t s = 0; s < m; ++s) {
    double x = col[s * n];
    finite = finite && isfinite(x);
    sum += x;
  }
  const double mean = sum / (double)m;
  double ss = 0.0;
  for (int64_t s = 0; s < m; ++s) {
    double d = col[s * n] - mean;
    ss += d * d;
  }
  const bool is_bad = !finite || !(ss > 0.0);
  const double nrm = sqrt(ss);
  if (live) bad[g] = is_bad ? 1 : 0;

  for (int64_t s0 = 0; s0 < ldk; s0 += 32) {
    for (int r = 0; r < 32; ++r) {
      int64_t s = s0 + r;
      double z = 0.0;
      if (s < m && !is_bad) z = (col[s * n] - mean) / nrm;
      tile[w][r][lane] = z;
    }
    __syncwarp();
    for (int r = 0; r < 32; ++r) {  // gene g0 + r, samples s0 + lane
      int64_t gg = g0 + r, s = s0 + lane;
      if (gg < n && s < ldk) Z[gg * ldk + s] = tile[w][lane][r];
    }
    __syncwarp();
  }
}
// -------------------------------------------------------------------------------------------
// K2  fused soft-threshold sweep.  The work is the CT x RT grid of 128 x 128 correlation tiles
// (CT column tiles of the genes whose connectivity is wanted, RT row tiles of all genes), taken
// column-major and cut into one contiguous, equally long range per CTA (persistent, one CTA per SM:
// no tail wave).  For each tile the CTA computes Z_i Z_j^T on the DMMA pipe and, while the tile is
// still in registers, evaluates clip -> similarity transform -> diag <- 1 -> NaN skip -> the running
// product chain for all P powers, reducing over rows.  The n x n matrix is never written.
// Column sums accumulate in shared memory with exactly one owner lane per (warp row, power, column);
// whenever the range crosses into the next column tile the sums are flushed to a partial slot, and
// sweep_reduce_kernel adds the slots of a column tile in CTA order: deterministic, no atomics.
//
// Epilogue shape: a quarter of the warp's 64 x 32 sub-tile (8 columns = 2 per thread, 8 rows each) is
// processed at a time; the transformed similarities overwrite the accumulator registers, the 16
// running products live in registers, and per power the 2 column sums go through tree adds and
// three shuffle rounds together, so their latencies overlap.
// -------------------------------------------------------------------------------------------
constexpr int kSweepStages = 3;
constexpr int kSweepMaxP = 32;

struct SweepParams {
  const double* Z;
  const uint8_t* bad;
  int64_t n, ldk;
  int64_t col0, ncols;  // genes whose connectivity this launch produces
  int64_t RT, T;        // row tiles, total tiles (CT * RT)
  int P;
  int net_type;
  int max_seg;          // partial slots per CTA
  double* partial;      // [gridDim.x][max_seg][P][128]
  double step[kSweepMaxP];
  int istep[kSweepMaxP];
};

__device__ __forceinline__ void sweep_range(int64_t T, int64_t G, int64_t b, int64_t& t0, int64_t& t1) {
  t0 = T * b / G;
  t1 = T * (b + 1) / G;
}

__global__ void __launch_bounds__(kGemmThreads, 1) sweep_kernel(const SweepParams prm) {
  extern __shared__ __align__(16) double smem[];
  double* colacc = smem + kSweepStages * kStageDoubles;  // [2][P][128]
  __shared__ uint8_t bad_col[kBN], bad_row[kBM];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
  const int P = prm.P;
  const int64_t n = prm.n;
  const int64_t jend = prm.col0 + prm.ncols;
  int64_t t0, t1;
  sweep_range(prm.T, gridDim.x, blockIdx.x, t0, t1);
  if (t0 >= t1) return;
  const int64_t jt_first = t0 / prm.RT;
  int64_t cur_jt = -1;

  for (int64_t t = t0; t < t1; ++t) {
    const int64_t jt = t / prm.RT, it = t - jt * prm.RT;
    const int64_t j0 = prm.col0 + jt * kBN, i0 = it * kBM;
    if (jt != cur_jt) {  // entering a new column tile: flush the previous one, reset the sums
      __syncthreads();
      if (cur_jt >= 0) {
        double* dst = prm.partial + ((int64_t)blockIdx.x * prm.max_seg + (cur_jt - jt_first)) * P * kBN;
        for (int i = threadIdx.x; i < P * kBN; i += kGemmThreads) dst[i] = colacc[i] + colacc[P * kBN + i];
        __syncthreads();
      }
      for (int i = threadIdx.x; i < 2 * P * kBN; i += kGemmThreads) colacc[i] = 0.0;
      if (threadIdx.x < kBN) {
        int64_t j = j0 + threadIdx.x;
        bad_col[threadIdx.x] = (j < jend) ? prm.bad[j] : 1;
      }
      cur_jt = jt;
    }
    if (threadIdx.x < kBM) {
      int64_t i = i0 + threadIdx.x;
      bad_row[threadIdx.x] = (i < n) ? prm.bad[i] : 1;
    }
    AccF64 acc;
    acc.clear();
    // rows of the tile = genes i (summed over), columns = genes j (kept)
    gemm_f64_tile<kSweepStages>(acc, smem, prm.Z, prm.Z, prm.ldk, prm.ldk, i0, n, j0, n);
    // (gemm_f64_tile ends with a __syncthreads, so bad_row / bad_col / the zeroed sums are visible)

#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double cur[16];
      // element e = (cb * 4 + mi) * 2 + ch  ->  acc.v[mi][ni][ch * 2 + cb]
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        const int col = wn * 32 + ni * 8 + 2 * tig + cb;
        const int64_t gj = j0 + col;
        const bool cbad = bad_col[col];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const int row = wm * 64 + mi * 16 + g + ch * 8;
            const int64_t gi = i0 + row;
            double c = acc.v[mi][ni][ch * 2 + cb];
            c = fmin(1.0, fmax(-1.0, c));                // np.corrcoef clips to [-1, 1]
            double sv = net_transform(c, prm.net_type);  // wgcna.py:884-889
            double c0 = 1.0;
            if (gi == gj) sv = 1.0;                      // wgcna.py:897 (even for a NaN gene)
            else if (cbad || bad_row[row]) { sv = 0.0; c0 = 0.0; }  // NaN: skipped by nansum :915
            if (gi >= n) { sv = 0.0; c0 = 0.0; }
            acc.v[mi][ni][ch * 2 + cb] = sv;
            cur[(cb * 4 + mi) * 2 + ch] = c0;
          }
      }
#pragma unroll 1
      for (int p = 0; p < P; ++p) {
        const int ist = prm.istep[p];
        // wgcna.py:914  corxCur = corxPrev * corx ** step
        if (ist == 1) {
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
              for (int ch = 0; ch < 2; ++ch) cur[(cb * 4 + mi) * 2 + ch] *= acc.v[mi][ni][ch * 2 + cb];
        } else {
          const double st = prm.step[p];
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
              for (int ch = 0; ch < 2; ++ch)
                cur[(cb * 4 + mi) * 2 + ch] *= step_pow(acc.v[mi][ni][ch * 2 + cb], st, ist);
        }
        double v[2];
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const double* q = cur + cb * 8;
          v[cb] = ((q[0] + q[1]) + (q[2] + q[3])) + ((q[4] + q[5]) + (q[6] + q[7]));
        }
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1)
#pragma unroll
          for (int cb = 0; cb < 2; ++cb) v[cb] += __shfl_xor_sync(0xffffffffu, v[cb], o);
        if (g == 0) {
#pragma unroll
          for (int cb = 0; cb < 2; ++cb)
            colacc[(wm * P + p) * kBN + wn * 32 + ni * 8 + 2 * tig + cb] += v[cb];
        }
      }
    }
  }
  __syncthreads();
  {
    double* dst = prm.partial + ((int64_t)blockIdx.x * prm.max_seg + (cur_jt - jt_first)) * P * kBN;
    for (int i = threadIdx.x; i < P * kBN; i += kGemmThreads) dst[i] = colacc[i] + colacc[P * kBN + i];
  }
}

// k[p][j] = sum over the CTAs whose range covers column tile jt of their slot - 1   (wgcna.py:915)
__global__ void sweep_reduce_kernel(const double* __restrict__ partial, int G, int max_seg, int P,
                                    int64_t RT, int64_t T, int64_t ncols, double* __restrict__ k) {
  const int64_t jt = blockIdx.x;
  const int col = threadIdx.x;  // 128
  const int p = blockIdx.y;
  const int64_t j = jt * kBN + col;
  if (j >= ncols) return;
  const int64_t lo = jt * RT, hi = lo + RT;  // tiles of this column tile
  double s = 0.0;
  for (int64_t b = lo * G / T; b < G; ++b) {
    int64_t t0, t1;
    sweep_range(T, G, b, t0, t1);
    if (t0 >= hi) break;
    if (t1 <= lo || t0 >= t1) continue;
    const int64_t seg = jt - t0 / RT;
    s += partial[(((int64_t)b * max_seg + seg) * P + p) * kBN + col];
  }
  k[(int64_t)p * ncols + j] = s - 1.0;
}

// -------------------------------------------------------------------------------------------
// K1 + K3  adjacency: correlation tile on the DMMA pipe, epilogue clip -> transform -> s ** power
// (wgcna.py:1097-1111).  Rows of a NaN gene are NaN, as np.corrcoef makes them; the diagonal is not
// forced (the reference leaves it at corrcoef's value).  Output rows [row0, row0 + nrows).
// Whole matrix (row0 = 0, nrows = n): only tiles on and above the diagonal are computed, every
// off-diagonal tile is stored twice (cor is symmetric: the same dot product), which also makes the
// result exactly symmetric -- checkAdjMat's 1e-12 test (wgcna.py:1135) passes with margin.
// Small integer powers are evaluated by repeated squaring (<= 2 ulp from libm's pow).
// -------------------------------------------------------------------------------------------
constexpr int kAdjStages = 4;

__device__ __forceinline__ double int_pow(double s, int e) {
  double r = 1.0, b = s;
#pragma unroll 1
  while (e) {
    if (e & 1) r *= b;
    b *= b;
    e >>= 1;
  }
  return r;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
adjacency_kernel(const double* __restrict__ Z, const uint8_t* __restrict__ bad, int64_t n,
                 int64_t ldk, int net_type, double power, int ipower, int64_t row0, int64_t nrows,
                 int symmetric, int64_t tiles_n, double* __restrict__ A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
  int64_t ti, tj;
  if (symmetric) {  // blockIdx.x enumerates the upper triangle row by row
    int64_t t = blockIdx.x, T = tiles_n;
    ti = (int64_t)floor((2.0 * T + 1.0 - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)t)) / 2.0);
    while (ti > 0 && ti * T - ti * (ti - 1) / 2 > t) --ti;
    while ((ti + 1) * T - (ti + 1) * ti / 2 <= t) ++ti;
    tj = ti + (t - (ti * T - ti * (ti - 1) / 2));
  } else {
    ti = blockIdx.x / tiles_n;
    tj = blockIdx.x % tiles_n;
  }
  const int64_t i0 = row0 + ti * kBM, j0 = tj * kBN;
  const int64_t iend = row0 + nrows;
  const bool mirror = symmetric && ti != tj;

  AccF64 acc;
  acc.clear();
  gemm_f64_tile<kAdjStages>(acc, smem, Z, Z, ldk, ldk, i0, n, j0, n);

  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int64_t gi = i0 + wm * 64 + mi * 16 + g + ch * 8;
      if (gi >= iend) continue;
      const bool rbad = bad[gi];
      double* out = A + (gi - row0) * n;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int64_t gj = j0 + wn * 32 + ni * 8 + 2 * tig + cb;
          if (gj >= n) continue;
          double c = acc.v[mi][ni][ch * 2 + cb];
          c = fmin(1.0, fmax(-1.0, c));
          double s = net_transform(c, net_type);
          double a;
          if (ipower > 0) a = int_pow(s, ipower);
          else a = pow(s, power);
          if (rbad || bad[gj]) a = qnan;
          out[gj] = a;
          if (mirror) A[gj * n + gi] = a;
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// Cross-correlation of two standardised gene sets, C[i, j] = cor(x_i, y_j): the n x M block that
// CalculateSignedKME takes out of np.corrcoef(datExpr.T, datME.T) (wgcna.py:3486-3489) without forming
// the (n + M)^2 matrix around it.  Same DMMA tile engine, epilogue clip + NaN rows/columns.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGemmThreads, 1)
cross_cor_kernel(const double* __restrict__ Zx, const uint8_t* __restrict__ badx, int64_t n,
                 const double* __restrict__ Zy, const uint8_t* __restrict__ bady, int64_t M, int64_t ldk,
                 double* __restrict__ C) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
  const int64_t i0 = (int64_t)blockIdx.y * kBM, j0 = (int64_t)blockIdx.x * kBN;
  AccF64 acc;
  acc.clear();
  gemm_f64_tile<kAdjStages>(acc, smem, Zx, Zy, ldk, ldk, i0, n, j0, M);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int64_t gi = i0 + wm * 64 + mi * 16 + g + ch * 8;
      if (gi >= n) continue;
      const bool rbad = badx[gi];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni)
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int64_t gj = j0 + wn * 32 + ni * 8 + 2 * tig + cb;
          if (gj >= M) continue;
          double c = fmin(1.0, fmax(-1.0, acc.v[mi][ni][ch * 2 + cb]));
          C[gi * M + gj] = (rbad || bady[gj]) ? qnan : c;
        }
    }
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_dev_cross_correlation(const double* dZx, const uint8_t* d_badx, int64_t n,
                                           const double* dZy, const uint8_t* d_bady, int64_t M, int64_t m,
                                           double* dC, int device, void* stream) {
  if (!dZx || !d_badx || !dZy || !d_bady || !dC || n < 1 || M < 1 || m < 1)
    return fail(B200W_E_INVALID, "cross_correlation: bad args");
  if (int e = ensure_device(device)) return e;
  const size_t smem = gemm_smem_bytes(kAdjStages);
  B200W_CUDA(cudaFuncSetAttribute(cross_cor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t rt = (n + kBM - 1) / kBM, ct = (M + kBN - 1) / kBN;
  if (rt > 65535) return fail(B200W_E_UNSUPPORTED, "cross_correlation: too many row tiles");
  dim3 grid((unsigned)ct, (unsigned)rt);
  cross_cor_kernel<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(dZx, d_badx, n, dZy, d_bady, M,
                                                                      b200w_padded_k(m), dC);
  B200W_LAUNCHED();
  return 0;
}

extern "C" int64_t b200w_padded_k(int64_t m) { return round_up(m, kBK); }

extern "C" int b200w_dev_standardize(const double* dX, int64_t m, int64_t n, double* dZ,
                                     uint8_t* d_bad, int device, void* stream) {
  if (!dX || !dZ || !d_bad || m < 1 || n < 1) return fail(B200W_E_INVALID, "standardize: bad args");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t ldk = b200w_padded_k(m);
  int64_t warps = (n + 31) / 32;
  unsigned blocks = (unsigned)((warps + kStdWarps - 1) / kStdWarps);
  standardize_kernel<<<blocks, kStdWarps * 32, 0, st>>>(dX, m, n, ldk, dZ, d_bad);
  B200W_LAUNCHED();
  return 0;
}

extern "C" int b200w_dev_sweep(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                               const double* h_steps, int P, int net_type, int64_t col0,
                               int64_t ncols, double* d_k, int device, void* stream) {
  if (!dZ || !d_bad || !h_steps || !d_k || P < 1 || n < 1 || col0 < 0 || ncols < 1 ||
      col0 + ncols > n)
    return fail(B200W_E_INVALID, "sweep: bad args");
  if (net_type < 0 || net_type > 2) return fail(B200W_E_INVALID, "sweep: bad network type");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldk = b200w_padded_k(m);
  const int64_t CT = (ncols + kBN - 1) / kBN, RT = (n + kBM - 1) / kBM, T = CT * RT;
  int sms = 0;
  B200W_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int G = (int)(T < sms ? T : sms);  // persistent: one CTA per SM, equal tile ranges
  const int max_seg = (int)((T / G + 1 + RT - 1) / RT + 1);

  const size_t smem_max = gemm_smem_bytes(kSweepStages) + (size_t)2 * kSweepMaxP * kBN * 8;
  B200W_CUDA(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_max));
  for (int p0 = 0; p0 < P; p0 += kSweepMaxP) {
    // the chain restarts from 1 for each chunk of 32 powers, so the first step of a later chunk is
    // the whole power reached so far (pow() with a large exponent: the value the reference's chain
    // approximates to ~1e-15).
    const int Pc = (P - p0 < kSweepMaxP) ? P - p0 : kSweepMaxP;
    SweepParams prm;
    prm.Z = dZ; prm.bad = d_bad; prm.n = n; prm.ldk = ldk; prm.col0 = col0; prm.ncols = ncols;
    prm.RT = RT; prm.T = T; prm.P = Pc; prm.net_type = net_type; prm.max_seg = max_seg;
    double base = 0.0;
    for (int q = 0; q < p0; ++q) base += h_steps[q];
    for (int q = 0; q < Pc; ++q) {
      double s = h_steps[p0 + q] + (q == 0 ? base : 0.0);
      prm.step[q] = s;
      prm.istep[q] = (s == (double)(int)s && s >= 1.0 && s <= 4.0) ? (int)s : 0;
    }
    Scratch part;
    B200W_CUDA(part.alloc((size_t)G * max_seg * Pc * kBN * sizeof(double), st));
    prm.partial = part.as<double>();
    const size_t smem = gemm_smem_bytes(kSweepStages) + (size_t)2 * Pc * kBN * 8;
    sweep_kernel<<<G, kGemmThreads, smem, st>>>(prm);
    B200W_LAUNCHED();
    dim3 rgrid((unsigned)CT, (unsigned)Pc);
    sweep_reduce_kernel<<<rgrid, kBN, 0, st>>>(prm.partial, G, max_seg, Pc, RT, T, ncols,
                                               d_k + (int64_t)p0 * ncols);
    B200W_LAUNCHED();
  }
  return 0;
}

extern "C" int b200w_dev_adjacency(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                                   int net_type, double power, int64_t row0, int64_t nrows,
                                   double* dA, int device, void* stream) {
  if (!dZ || !d_bad || !dA || n < 1 || row0 < 0 || nrows < 1 || row0 + nrows > n)
    return fail(B200W_E_INVALID, "adjacency: bad args");
  if (net_type < 0 || net_type > 2) return fail(B200W_E_INVALID, "adjacency: bad network type");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = gemm_smem_bytes(kAdjStages);
  B200W_CUDA(cudaFuncSetAttribute(adjacency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  int ipower = (power == (double)(int)power && power >= 1.0 && power <= 64.0) ? (int)power : 0;
  const int64_t row_tiles = (nrows + kBM - 1) / kBM, col_tiles = (n + kBN - 1) / kBN;
  const int symmetric = (row0 == 0 && nrows == n) ? 1 : 0;
  const int64_t blocks = symmetric ? col_tiles * (col_tiles + 1) / 2 : row_tiles * col_tiles;
  if (blocks > 2147483647LL) return fail(B200W_E_UNSUPPORTED, "adjacency: too many tiles");
  adjacency_kernel<<<(unsigned)blocks, kGemmThreads, smem, st>>>(dZ, d_bad, n, b200w_padded_k(m),
                                                                  net_type, power, ipower, row0, nrows,
                                                                  symmetric, col_tiles, dA);
  B200W_LAUNCHED();
  return 0;
}
