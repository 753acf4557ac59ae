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
// K2  fused soft-threshold sweep.  CTA (jt, rc) owns the 128 genes of column tile jt and the row
// tiles it = rc, rc + R, ... ; for each it computes the 128 x 128 correlation tile on the DMMA
// pipe and, while it is still in registers, evaluates the similarity transform and the running
// product chain for all P powers, reducing over rows.  The n x n matrix is never written.
// Column sums accumulate in shared memory with exactly one owner lane per (warp row, power, column),
// and the R partial results are summed in fixed order by sweep_reduce_kernel: the result is
// deterministic and independent of the launch geometry of other CTAs.
// -------------------------------------------------------------------------------------------
constexpr int kSweepStages = 3;
constexpr int kSweepMaxP = 32;

struct SweepParams {
  const double* Z;
  const uint8_t* bad;
  int64_t n, ldk;
  int64_t col0, ncols;  // genes whose connectivity this launch produces
  int P;
  int net_type;
  int R;             // row chunks
  double* partial;   // [R][P][ncols_pad]
  int64_t ncols_pad;
  double step[kSweepMaxP];
  int istep[kSweepMaxP];
};

__global__ void __launch_bounds__(kGemmThreads, 1) sweep_kernel(const SweepParams prm) {
  extern __shared__ __align__(16) double smem[];
  double* colacc = smem + kSweepStages * kStageDoubles;  // [2][P][128]
  __shared__ uint8_t bad_col[kBN], bad_row[kBM];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
  const int P = prm.P;
  const int64_t n = prm.n;
  const int64_t j0 = prm.col0 + (int64_t)blockIdx.x * kBN;  // first gene of the column tile
  const int64_t jend = prm.col0 + prm.ncols;

  for (int i = threadIdx.x; i < 2 * P * kBN; i += kGemmThreads) colacc[i] = 0.0;
  if (threadIdx.x < kBN) {
    int64_t j = j0 + threadIdx.x;
    bad_col[threadIdx.x] = (j < jend) ? prm.bad[j] : 1;
  }
  __syncthreads();

  const int64_t n_row_tiles = (n + kBM - 1) / kBM;
  for (int64_t it = blockIdx.y; it < n_row_tiles; it += prm.R) {
    const int64_t i0 = it * kBM;
    if (threadIdx.x < kBM) {
      int64_t i = i0 + threadIdx.x;
      bad_row[threadIdx.x] = (i < n) ? prm.bad[i] : 1;
    }
    AccF64 acc;
    acc.clear();
    // rows of the tile = genes i (summed over), columns = genes j (kept)
    gemm_f64_tile<kSweepStages>(acc, smem, prm.Z, prm.Z, prm.ldk, prm.ldk, i0, n, j0, n);
    // (gemm_f64_tile ends with a __syncthreads, so bad_row is visible)

#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        const int col = wn * 32 + ni * 8 + 2 * tig + cb;
        const int64_t gj = j0 + col;
        const bool cbad = bad_col[col];
        double s[8], cur[8];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const int row = wm * 64 + mi * 16 + g + ch * 8;
            const int64_t gi = i0 + row;
            double c = acc.v[mi][ni][ch * 2 + cb];
            c = fmin(1.0, fmax(-1.0, c));                  // np.corrcoef clips to [-1, 1]
            double sv = net_transform(c, prm.net_type);    // wgcna.py:884-889
            double c0 = 1.0;
            if (gi == gj) sv = 1.0;                        // wgcna.py:897 (even for a NaN gene)
            else if (cbad || bad_row[row]) { sv = 0.0; c0 = 0.0; }  // NaN: skipped by nansum :915
            if (gi >= n) { sv = 0.0; c0 = 0.0; }
            s[mi * 2 + ch] = sv;
            cur[mi * 2 + ch] = c0;
          }
        }
        for (int p = 0; p < P; ++p) {
          const double st = prm.step[p];
          const int ist = prm.istep[p];
          double v = 0.0;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            cur[r] *= step_pow(s[r], st, ist);  // wgcna.py:914  corxCur = corxPrev * corx**step
            v += cur[r];
          }
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) colacc[(wm * P + p) * kBN + col] += v;
        }
      }
    }
    __syncthreads();  // bad_row is rewritten by the next iteration
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P * kBN; i += kGemmThreads) {
    int p = i / kBN, col = i % kBN;
    int64_t jj = (int64_t)blockIdx.x * kBN + col;
    if (jj < prm.ncols)
      prm.partial[((int64_t)blockIdx.y * P + p) * prm.ncols_pad + jj] =
          colacc[p * kBN + col] + colacc[(P + p) * kBN + col];
  }
}

// k[p][j] = sum_r partial[r][p][j] - 1      (wgcna.py:915  nansum(...) - 1)
__global__ void sweep_reduce_kernel(const double* __restrict__ partial, int R, int P,
                                    int64_t ncols, int64_t ncols_pad, double* __restrict__ k) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int p = blockIdx.y;
  if (j >= ncols) return;
  double s = 0.0;
  for (int r = 0; r < R; ++r) s += partial[((int64_t)r * P + p) * ncols_pad + j];
  k[(int64_t)p * ncols + j] = s - 1.0;
}

// -------------------------------------------------------------------------------------------
// K1 + K3  adjacency: correlation tile on the DMMA pipe, epilogue clip -> transform -> pow(., power)
// (wgcna.py:1097-1111).  Rows of a NaN gene are NaN, as np.corrcoef makes them; the diagonal is not
// forced (the reference leaves it at corrcoef's value).  Output rows [row0, row0 + nrows).
// -------------------------------------------------------------------------------------------
constexpr int kAdjStages = 4;

__global__ void __launch_bounds__(kGemmThreads, 1)
adjacency_kernel(const double* __restrict__ Z, const uint8_t* __restrict__ bad, int64_t n,
                 int64_t ldk, int net_type, double power, int ipower, int64_t row0, int64_t nrows,
                 double* __restrict__ A) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
  const int64_t i0 = row0 + (int64_t)blockIdx.y * kBM, j0 = (int64_t)blockIdx.x * kBN;
  const int64_t iend = row0 + nrows;

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
          if (ipower == 1) a = s;
          else if (ipower == 2) a = s * s;
          else a = pow(s, power);
          if (rbad || bad[gj]) a = qnan;
          out[gj] = a;
        }
      }
    }
  }
}

}  // namespace b200w

using namespace b200w;

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
  const int64_t col_tiles = (ncols + kBN - 1) / kBN, row_tiles = (n + kBM - 1) / kBM;
  const int64_t ncols_pad = col_tiles * kBN;
  // row chunks: enough CTAs for >= 2 waves of 148 SMs, never more than there are row tiles
  int R = (int)((2 * 148 + col_tiles - 1) / col_tiles);
  if (R > row_tiles) R = (int)row_tiles;
  if (R < 1) R = 1;
  if (R > 65535) R = 65535;

  static bool attr_set = false;
  const size_t smem_max = gemm_smem_bytes(kSweepStages) + (size_t)2 * kSweepMaxP * kBN * 8;
  if (!attr_set) {
    B200W_CUDA(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_max));
    attr_set = true;
  }
  for (int p0 = 0; p0 < P; p0 += kSweepMaxP) {
    // the chain restarts from 1 for each chunk, so the first step of a later chunk is the whole
    // power reached so far; that needs pow() with a large exponent, which is the same value the
    // reference's chain approximates to ~1e-15.
    const int Pc = (P - p0 < kSweepMaxP) ? P - p0 : kSweepMaxP;
    SweepParams prm;
    prm.Z = dZ; prm.bad = d_bad; prm.n = n; prm.ldk = ldk; prm.col0 = col0; prm.ncols = ncols;
    prm.P = Pc; prm.net_type = net_type; prm.R = R; prm.ncols_pad = ncols_pad;
    double base = 0.0;
    for (int q = 0; q < p0; ++q) base += h_steps[q];
    for (int q = 0; q < Pc; ++q) {
      double s = h_steps[p0 + q] + (q == 0 ? base : 0.0);
      prm.step[q] = s;
      prm.istep[q] = (s == (double)(int)s && s >= 1.0 && s <= 4.0) ? (int)s : 0;
    }
    Scratch part;
    B200W_CUDA(part.alloc((size_t)R * Pc * ncols_pad * sizeof(double), st));
    prm.partial = part.as<double>();
    const size_t smem = gemm_smem_bytes(kSweepStages) + (size_t)2 * Pc * kBN * 8;
    dim3 grid((unsigned)col_tiles, (unsigned)R);
    sweep_kernel<<<grid, kGemmThreads, smem, st>>>(prm);
    B200W_LAUNCHED();
    dim3 rgrid((unsigned)((ncols + 255) / 256), (unsigned)Pc);
    sweep_reduce_kernel<<<rgrid, 256, 0, st>>>(prm.partial, R, Pc, ncols, ncols_pad,
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
  static bool attr_set = false;
  const size_t smem = gemm_smem_bytes(kAdjStages);
  if (!attr_set) {
    B200W_CUDA(cudaFuncSetAttribute(adjacency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    attr_set = true;
  }
  int ipower = (power == 1.0) ? 1 : (power == 2.0) ? 2 : 0;
  int64_t row_tiles = (nrows + kBM - 1) / kBM, col_tiles = (n + kBN - 1) / kBN;
  if (row_tiles > 65535) return fail(B200W_E_UNSUPPORTED, "adjacency: too many row tiles");
  dim3 grid((unsigned)col_tiles, (unsigned)row_tiles);
  adjacency_kernel<<<grid, kGemmThreads, smem, st>>>(dZ, d_bad, n, b200w_padded_k(m), net_type,
                                                     power, ipower, row0, nrows, dA);
  B200W_LAUNCHED();
  return 0;
}
