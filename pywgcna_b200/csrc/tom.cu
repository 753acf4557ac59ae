// Topological-overlap side of the hot path: checkAdjMat reductions, operand preparation,
// the FP64 (DMMA) contraction + TOM epilogue, intramodular connectivity.
// Reference: PyWGCNA/wgcna.py:1114-1138 (checkAdjMat), :1141-1157 (TomSimilarityFromAdj),
// :1160-1193 (TOMsimilarity), :3403-3450 (intramodularConnectivity).
#include <cuda.h>

#include "gemm_f64.cuh"
#include "tom_common.cuh"

namespace b200w {

// -------------------------------------------------------------------------------------------
// checkAdjMat reductions (wgcna.py:1135-1138): max|A - A^T|, min, max, NaN count in one pass, plus
// max|A - A^T| of the matrix after nan_to_num (what TomSimilarityFromAdj really gets, :1185 -> :1189).
// Each CTA handles one 32 x 32 tile pair (bi <= bj): it reads tile (bi,bj) and tile (bj,bi), both
// row-wise (coalesced), transposes the second through shared memory.  Per-CTA partials are reduced
// by a second kernel in fixed order.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
check_adj_kernel(const double* __restrict__ A, int64_t n, int64_t tiles, double* __restrict__ part) {
  __shared__ double tb[32][33];
  __shared__ double red[6][8];
  // triangular tile index -> (bi, bj), bi <= bj
  int64_t t = blockIdx.x;
  int64_t bi = (int64_t)floor((2.0 * tiles + 1.0 - sqrt((2.0 * tiles + 1.0) * (2.0 * tiles + 1.0) - 8.0 * (double)t)) / 2.0);
  while (bi > 0 && bi * tiles - bi * (bi - 1) / 2 > t) --bi;
  while ((bi + 1) * tiles - (bi + 1) * bi / 2 <= t) ++bi;
  int64_t bj = bi + (t - (bi * tiles - bi * (bi - 1) / 2));
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8

  double asym = 0.0, mn = INFINITY, mx = -INFINITY, nans = 0.0, asymc = 0.0, nanasym = 0.0;
  for (int r = ty; r < 32; r += 8) {  // tile (bj, bi) rows
    int64_t i = bj * 32 + r, j = bi * 32 + tx;
    tb[r][tx] = (i < n && j < n) ? A[i * n + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int64_t i = bi * 32 + r, j = bj * 32 + tx;
    if (i < n && j < n) {
      double a = A[i * n + j], b = tb[tx][r];
      asym = fmax(asym, fabs(a - b));  // fmax drops a NaN operand
      asymc = fmax(asymc, fabs((isnan(a) ? 0.0 : a) - (isnan(b) ? 0.0 : b)));  // after nan_to_num (wgcna.py:1185)
      if (isnan(a) != isnan(b)) nanasym += 1.0;  // NaN pattern not symmetric
      if (isnan(a)) nans += 1.0; else { mn = fmin(mn, a); mx = fmax(mx, a); }
      if (bi != bj) { if (isnan(b)) nans += 1.0; else { mn = fmin(mn, b); mx = fmax(mx, b); } }
    }
  }
  for (int o = 16; o; o >>= 1) {
    asym = fmax(asym, __shfl_xor_sync(0xffffffffu, asym, o));
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    nans += __shfl_xor_sync(0xffffffffu, nans, o);
    asymc = fmax(asymc, __shfl_xor_sync(0xffffffffu, asymc, o));
    nanasym += __shfl_xor_sync(0xffffffffu, nanasym, o);
  }
  if (tx == 0) {
    red[0][ty] = asym; red[1][ty] = mn; red[2][ty] = mx; red[3][ty] = nans; red[4][ty] = asymc;
    red[5][ty] = nanasym;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      asym = fmax(asym, red[0][w]); mn = fmin(mn, red[1][w]); mx = fmax(mx, red[2][w]);
      nans += red[3][w]; asymc = fmax(asymc, red[4][w]); nanasym += red[5][w];
    }
    double* o = part + (int64_t)blockIdx.x * 6;
    o[0] = asym; o[1] = mn; o[2] = mx; o[3] = nans; o[4] = asymc; o[5] = nanasym;
  }
}

__global__ void check_adj_reduce_kernel(const double* __restrict__ part, int64_t nb,
                                        double* __restrict__ stats) {
  __shared__ double red[6][256];
  double asym = 0.0, mn = INFINITY, mx = -INFINITY, nans = 0.0, asymc = 0.0, nanasym = 0.0;
  for (int64_t b = threadIdx.x; b < nb; b += blockDim.x) {
    asym = fmax(asym, part[b * 6]); mn = fmin(mn, part[b * 6 + 1]);
    mx = fmax(mx, part[b * 6 + 2]); nans += part[b * 6 + 3]; asymc = fmax(asymc, part[b * 6 + 4]);
    nanasym += part[b * 6 + 5];
  }
  red[0][threadIdx.x] = asym; red[1][threadIdx.x] = mn; red[2][threadIdx.x] = mx;
  red[3][threadIdx.x] = nans; red[4][threadIdx.x] = asymc; red[5][threadIdx.x] = nanasym;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)blockDim.x; ++i) {
      asym = fmax(asym, red[0][i]); mn = fmin(mn, red[1][i]); mx = fmax(mx, red[2][i]);
      nans += red[3][i]; asymc = fmax(asymc, red[4][i]); nanasym += red[5][i];
    }
    stats[0] = asym; stats[1] = mn; stats[2] = mx; stats[3] = nans; stats[4] = asymc; stats[5] = nanasym;
  }
}

// -------------------------------------------------------------------------------------------
// Operand preparation, FP64 flavour: one CTA per row.  Ac[i, :] = A[i, :] with NaN -> 0
// (wgcna.py:1185) and the diagonal zeroed (:1143), zero padded to ldn; k_i = sum_j Ac[i, j] (:1146).
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tom_prepare_f64_kernel(const double* __restrict__ A_rows, int64_t n, int64_t ldn, int64_t row0,
                       double* __restrict__ Ac, double* __restrict__ k, int nan_mode) {
  const int64_t i = row0 + blockIdx.x;
  const double* src = A_rows + (int64_t)blockIdx.x * n;
  double* dst = Ac + i * ldn;
  double s = 0.0, nn = 0.0;
  for (int64_t j = threadIdx.x; j < ldn; j += 256) {
    double a = 0.0;
    if (j < n) {
      if (i != j && isnan(src[j])) nn += 1.0;
      a = tom_clean(src[j], i, j);
    }
    dst[j] = a;
    s += a;
  }
  s = block_sum_256(s);
  nn = block_sum_256(nn);
  // nan_mode 1 (TomSimilarityFromAdj without the nan_to_num of TOMsimilarity): a NaN in the row makes its
  // plain numpy sum NaN (wgcna.py:1146-1147) and the epilogue propagates it
  if (threadIdx.x == 0) k[i] = (nan_mode && nn > 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : s;
}

// At[j, i] = A[i, j]: the column operand of an asymmetric A (TomSimilarityFromAdj forms A @ A with the
// column sums k_j, wgcna.py:1145-1147; the kernels contract rows with rows)
__global__ void __launch_bounds__(256)
transpose_kernel(const double* __restrict__ A, int64_t n, double* __restrict__ At) {
  __shared__ double t[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t bi = (int64_t)blockIdx.y * 32, bj = (int64_t)blockIdx.x * 32;
  for (int r = ty; r < 32; r += 8)
    if (bi + r < n && bj + tx < n) t[r][tx] = A[(bi + r) * n + bj + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bj + r < n && bi + tx < n) At[(bj + r) * n + bi + tx] = t[tx][r];
}

// -------------------------------------------------------------------------------------------
// K4 (FP64 flavour): L = Ac[rows] * Ac^T on the DMMA pipe (A symmetric to 1e-12, checkAdjMat),
// epilogue (L + a) / (min(k_i, k_j) + 1 - a) etc. while the tile is in registers.
// 1-D grid, row tiles grouped by 8 so that concurrently resident CTAs share operand panels in L2.
// -------------------------------------------------------------------------------------------
constexpr int kTomStages = 4;

__global__ void __launch_bounds__(kGemmThreads, 1)
tom_f64_kernel(const double* __restrict__ A_rows, const double* __restrict__ Ac,
               const double* __restrict__ Bc, const double* __restrict__ k,
               const double* __restrict__ k_col, int nan_mode, int64_t n, int64_t ldn, int64_t row0,
               int64_t nrows, int tom_type, int tom_denom, int out_mode, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, tig = lane & 3;
  const int64_t tiles_m = (nrows + kBM - 1) / kBM, tiles_n = (n + kBN - 1) / kBN;
  int64_t tm, tn;
  grouped_tile(blockIdx.x, tiles_m, tiles_n, 8, tm, tn);
  const int64_t i0 = row0 + tm * kBM, j0 = tn * kBN, iend = row0 + nrows;

  AccF64 acc;
  acc.clear();
  gemm_f64_tile<kTomStages>(acc, smem, Ac, Bc, ldn, ldn, i0, n, j0, n);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const int64_t gi = i0 + wm * 64 + mi * 16 + g + ch * 8;
      if (gi >= iend) continue;
      const double ki = k[gi];
      const double* arow = A_rows + (gi - row0) * n;
      double* orow = out + (gi - row0) * n;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int64_t gj = j0 + wn * 32 + ni * 8 + 2 * tig + cb;
          if (gj >= n) continue;
          const double a = tom_clean(arow[gj], gi, gj);
          const double kj = k_col[gj];
          double t = tom_value(acc.v[mi][ni][ch * 2 + cb], a, ki, kj, gi == gj, tom_type, tom_denom, out_mode);
          if (nan_mode && gi != gj && (isnan(ki) || isnan(kj) || isnan(arow[gj]))) t = qnan;
          if (!out_is_condensed(out_mode)) orow[gj] = t;
          else if (gj > gi) out[condensed_index(n, gi, gj)] = t;
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// intramodularConnectivity (wgcna.py:3431-3445): one CTA per row, HBM-bound segmented row sum.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
intramodular_kernel(const double* __restrict__ A, int64_t n, const int32_t* __restrict__ labels,
                    const uint8_t* __restrict__ small, double* __restrict__ k_total,
                    double* __restrict__ k_within) {
  const int64_t i = blockIdx.x;
  const int32_t li = labels[i];
  const double* row = A + i * n;
  double st = 0.0, sw = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += 256) {
    double a = tom_clean(row[j], i, j);
    st += a;
    if (labels[j] == li) sw += a;
  }
  st = block_sum_256(st);
  sw = block_sum_256(sw);
  if (threadIdx.x == 0) {
    k_total[i] = st;
    k_within[i] = small[li] ? 0.0 : sw;
  }
}

}  // namespace b200w

using namespace b200w;

extern "C" int64_t b200w_padded_n(int64_t n) { return round_up(n, 128); }

extern "C" int64_t b200w_tom_operand_bytes(int64_t n, int precision) {
  const int64_t ldn = b200w_padded_n(n);
  // rows are padded to a whole number of 128-row operand tiles as well, so that TMA boxes of the
  // int8 path never start outside the tensor
  const int64_t rows = round_up(n, 128);
  if (resolve_precision(precision) == B200W_PREC_F64) return rows * ldn * 8;
  return (int64_t)B200W_I8_SLICES * rows * ldn;
}

extern "C" int b200w_dev_check_adjacency(const double* dA, int64_t n, double* d_stats6, int device,
                                         void* stream) {
  double* const d_stats4 = d_stats6;
  if (!dA || !d_stats4 || n < 1) return fail(B200W_E_INVALID, "check_adjacency: bad args");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = (n + 31) / 32, nb = tiles * (tiles + 1) / 2;
  if (nb > 2147483647LL) return fail(B200W_E_UNSUPPORTED, "check_adjacency: n too large");
  Scratch part;
  B200W_CUDA(part.alloc((size_t)nb * 6 * sizeof(double), st));
  check_adj_kernel<<<(unsigned)nb, 256, 0, st>>>(dA, n, tiles, part.as<double>());
  B200W_LAUNCHED();
  check_adj_reduce_kernel<<<1, 256, 0, st>>>(part.as<double>(), nb, d_stats4);
  B200W_LAUNCHED();
  return 0;
}

int b200w_tom_prepare_i8(const double* dA_rows, int64_t n, int64_t row0, int64_t nrows,
                         void* d_operand, int64_t op_rows, double* d_scale, double* d_k,
                         cudaStream_t st, int nan_mode);
int b200w_tom_compute_i8(const double* dA_rows, const void* d_operand, int64_t op_rows,
                         const double* d_scale,
                         const double* d_k, int64_t n, int64_t row0, int64_t nrows, int tom_type,
                         int tom_denom, int out_mode, double* d_out, cudaStream_t st, int world,
                         int rank, int64_t per, double* const* peer_out, b200w::TomProgress* progress,
                         int w_min, const b200w::TomColumnOperand* col, const unsigned int* d_plane_flags = nullptr,
                         unsigned int plane_epoch = 0);

int b200w::resolve_precision(int precision) {
  if (precision == B200W_PREC_AUTO) return B200W_PREC_I8;
  return precision;
}

extern "C" int b200w_dev_tom_prepare(const double* dA_rows, int64_t n, int64_t row0, int64_t nrows,
                                     int precision, void* d_operand, int64_t op_rows,
                                     double* d_scale, double* d_k, int device, void* stream) {
  return b200w::tom_prepare_impl(dA_rows, n, row0, nrows, precision, d_operand, op_rows, d_scale, d_k, device,
                                 stream, 0);
}

int b200w::tom_transpose(const double* dA, int64_t n, double* dAt, cudaStream_t st) {
  const int64_t t = (n + 31) / 32;
  if (t > 65535) return fail(B200W_E_UNSUPPORTED, "tom: n too large for the transposed operand");
  transpose_kernel<<<dim3((unsigned)t, (unsigned)t), 256, 0, st>>>(dA, n, dAt);
  B200W_LAUNCHED();
  return 0;
}

int b200w::tom_prepare_impl(const double* dA_rows, int64_t n, int64_t row0, int64_t nrows, int precision,
                            void* d_operand, int64_t op_rows, double* d_scale, double* d_k, int device,
                            void* stream, int nan_mode) {
  if (!dA_rows || !d_operand || !d_k || n < 1 || row0 < 0 || nrows < 1 || row0 + nrows > n)
    return fail(B200W_E_INVALID, "tom_prepare: bad args");
  if (op_rows < round_up(n, 128) || op_rows % 128)
    return fail(B200W_E_INVALID, "tom_prepare: op_rows must be a multiple of 128 >= n");
  precision = resolve_precision(precision);
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == B200W_PREC_F64) {
    tom_prepare_f64_kernel<<<(unsigned)nrows, 256, 0, st>>>(dA_rows, n, b200w_padded_n(n), row0,
                                                            (double*)d_operand, d_k, nan_mode);
    B200W_LAUNCHED();
    return 0;
  }
  if (prec_is_i8(precision)) {
    if (!d_scale) return fail(B200W_E_INVALID, "tom_prepare: i8 needs d_scale");
    return b200w_tom_prepare_i8(dA_rows, n, row0, nrows, d_operand, op_rows, d_scale, d_k, st, nan_mode);
  }
  return fail(B200W_E_INVALID, "tom_prepare: unknown precision %d", precision);
}

extern "C" int b200w_dev_tom_compute(const double* dA_rows, const void* d_operand,
                                     int64_t op_rows, const double* d_scale, const double* d_k,
                                     int64_t n,
                                     int64_t row0, int64_t nrows, int tom_type, int tom_denom,
                                     int out_mode, int precision, double* d_out, int device,
                                     void* stream) {
  return b200w::tom_compute_impl(dA_rows, d_operand, op_rows, d_scale, d_k, n, row0, nrows, tom_type, tom_denom,
                                 out_mode, precision, d_out, device, stream, nullptr);
}

int b200w::tom_compute_impl(const double* dA_rows, const void* d_operand, int64_t op_rows, const double* d_scale,
                            const double* d_k, int64_t n, int64_t row0, int64_t nrows, int tom_type,
                            int tom_denom, int out_mode, int precision, double* d_out, int device, void* stream,
                            const TomColumnOperand* col) {
  if (!dA_rows || !d_operand || !d_k || !d_out || n < 1 || row0 < 0 || nrows < 1 ||
      row0 + nrows > n)
    return fail(B200W_E_INVALID, "tom_compute: bad args");
  if (op_rows < round_up(n, 128) || op_rows % 128)
    return fail(B200W_E_INVALID, "tom_compute: op_rows must be a multiple of 128 >= n");
  if (tom_type < 0 || tom_type > 1 || tom_denom < 0 || tom_denom > 1 || out_mode < 0 || out_mode > 4)
    return fail(B200W_E_INVALID, "tom_compute: bad enum");
  if (out_is_condensed(out_mode) && (row0 != 0 || nrows != n))
    return fail(B200W_E_INVALID, "tom_compute: condensed output needs the whole matrix (row0 = 0, nrows = n)");
  precision = resolve_precision(precision);
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == B200W_PREC_F64) {
    const size_t smem = gemm_smem_bytes(kTomStages);
    B200W_CUDA(cudaFuncSetAttribute(tom_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    const int64_t tiles = ((nrows + kBM - 1) / kBM) * ((n + kBN - 1) / kBN);
    tom_f64_kernel<<<(unsigned)tiles, kGemmThreads, smem, st>>>(
        dA_rows, (const double*)d_operand, col ? (const double*)col->operand : (const double*)d_operand, d_k,
        col ? col->k : d_k, col ? col->nan_mode : 0, n, b200w_padded_n(n), row0, nrows, tom_type, tom_denom,
        out_mode, d_out);
    B200W_LAUNCHED();
    return 0;
  }
  if (prec_is_i8(precision)) {
    if (!d_scale) return fail(B200W_E_INVALID, "tom_compute: i8 needs d_scale");
    return b200w_tom_compute_i8(dA_rows, d_operand, op_rows, d_scale, d_k, n, row0, nrows, tom_type,
                                tom_denom, out_mode, d_out, st, 1, 0, 0, nullptr, nullptr, prec_w_min(precision),
                                col);
  }
  return fail(B200W_E_INVALID, "tom_compute: unknown precision %d", precision);
}

static int tom_compute_dist_impl(const double* dA_rows, const void* d_operand, int64_t op_rows,
                                 const double* d_scale, const double* d_k, int64_t n, int world, int rank,
                                 int64_t per, int tom_type, int tom_denom, int out_mode, void* const* peer_out,
                                 const unsigned int* d_plane_flags, unsigned int plane_epoch, int device,
                                 void* stream) {
  if (!dA_rows || !d_operand || !d_scale || !d_k || !peer_out || n < 1 || world < 1 || world > 8 ||
      rank < 0 || rank >= world || per < 256 || per % 256)
    return fail(B200W_E_INVALID, "tom_compute_dist: bad args");
  if (tom_type < 0 || tom_type > 1 || tom_denom < 0 || tom_denom > 1 || out_mode < 0 || out_mode > 2)
    return fail(B200W_E_INVALID, "tom_compute_dist: bad enum");
  if (op_rows < round_up(n, 128) || op_rows % 128)
    return fail(B200W_E_INVALID, "tom_compute_dist: op_rows must be a multiple of 128 >= n");
  const int64_t row0 = (int64_t)rank * per;
  if (row0 >= n) return 0;  // a trailing rank without rows
  const int64_t nrows = (row0 + per <= n ? per : n - row0);
  if (int e = ensure_device(device)) return e;
  for (int r = 0; r < world; ++r)
    if (!peer_out[r] && (int64_t)r * per < n) return fail(B200W_E_INVALID, "tom_compute_dist: peer_out[%d] is NULL", r);
  return b200w_tom_compute_i8(dA_rows, d_operand, op_rows, d_scale, d_k, n, row0, nrows, tom_type,
                              tom_denom, out_mode, (double*)peer_out[rank], (cudaStream_t)stream, world,
                              rank, per, (double* const*)peer_out, nullptr, prec_w_min(B200W_PREC_I8), nullptr,
                              d_plane_flags, plane_epoch);
}

extern "C" int b200w_dev_tom_compute_dist(const double* dA_rows, const void* d_operand,
                                          int64_t op_rows, const double* d_scale, const double* d_k,
                                          int64_t n, int world, int rank, int64_t per, int tom_type,
                                          int tom_denom, int out_mode, void* const* peer_out,
                                          int device, void* stream) {
  return tom_compute_dist_impl(dA_rows, d_operand, op_rows, d_scale, d_k, n, world, rank, per, tom_type,
                               tom_denom, out_mode, peer_out, nullptr, 0u, device, stream);
}

extern "C" int b200w_dev_tom_compute_dist_gated(const double* dA_rows, const void* d_operand,
                                                int64_t op_rows, const double* d_scale, const double* d_k,
                                                int64_t n, int world, int rank, int64_t per, int tom_type,
                                                int tom_denom, int out_mode, void* const* peer_out,
                                                const uint32_t* d_plane_flags, uint32_t plane_epoch,
                                                int device, void* stream) {
  if (!d_plane_flags) return fail(B200W_E_INVALID, "tom_compute_dist_gated: bad args");
  return tom_compute_dist_impl(dA_rows, d_operand, op_rows, d_scale, d_k, n, world, rank, per, tom_type,
                               tom_denom, out_mode, peer_out, d_plane_flags, plane_epoch, device, stream);
}

// Pull the other ranks' operand rows (scale, k and the digit planes), peer by peer, out of their HBM with the copy
// engines, and publish each peer's arrival to the contraction kernel.
typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

extern "C" int b200w_dev_tom_pull_planes(void* d_local, void* const* peer_base, int world, int rank, int64_t per,
                                         int64_t op_rows, int64_t n, uint32_t* d_plane_flags, uint32_t epoch,
                                         int device, void* copy_stream) {
  if (!d_local || !peer_base || !d_plane_flags || world < 2 || world > 8 || rank < 0 || rank >= world ||
      per < 128 || per * world != op_rows)
    return fail(B200W_E_INVALID, "tom_pull_planes: bad args");
  if (int e = ensure_device(device)) return e;
  static StreamWriteValue32Fn write32 = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return (StreamWriteValue32Fn)sym;
  }();
  if (!write32) return fail(B200W_E_UNSUPPORTED, "tom_pull_planes: cuStreamWriteValue32 not available");
  cudaStream_t cs = (cudaStream_t)copy_stream;
  const int64_t ldn = b200w_padded_n(n);
  const int64_t plane = op_rows * ldn;
  const int64_t scale_off = b200w_tom_shared_scale_offset(op_rows, n), k_off = scale_off + op_rows * 8;
  auto pull = [&](int64_t off, int64_t bytes) -> cudaError_t {
    for (int q = 1; q < world; ++q) {  // start with the neighbour: the ranks do not all hit one peer first
      const int src = (rank + q) % world;
      if ((int64_t)src * per >= n) continue;  // a trailing rank without rows has nothing to give
      if (!peer_base[src]) return cudaErrorInvalidValue;
      cudaError_t e = cudaMemcpyAsync((char*)d_local + off + (int64_t)src * per * (bytes / per),
                                      (const char*)peer_base[src] + off + (int64_t)src * per * (bytes / per), bytes,
                                      cudaMemcpyDeviceToDevice, cs);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  for (int q = 1; q < world; ++q) {  // the order the contraction's tile list visits the peers in
    const int src = (rank + q) % world;
    if ((int64_t)src * per < n) {  // (a trailing rank without rows has nothing to give)
      if (!peer_base[src]) return fail(B200W_E_INVALID, "tom_pull_planes: peer_base[%d] is NULL", src);
      auto pull = [&](int64_t off, int64_t row_bytes) -> cudaError_t {
        const int64_t at = off + (int64_t)src * per * row_bytes;
        return cudaMemcpyAsync((char*)d_local + at, (const char*)peer_base[src] + at, (size_t)(per * row_bytes),
                               cudaMemcpyDeviceToDevice, cs);
      };
      B200W_CUDA(pull(scale_off, 8));
      B200W_CUDA(pull(k_off, 8));
      for (int s = B200W_I8_SLICES - 1; s >= 0; --s) B200W_CUDA(pull((int64_t)s * plane, ldn));
    }
    const CUresult r = write32((CUstream)cs, (CUdeviceptr)(d_plane_flags + src), epoch, 0);
    if (r != CUDA_SUCCESS) return fail(B200W_E_CUDA, "tom_pull_planes: cuStreamWriteValue32 failed (%d)", (int)r);
  }
  return 0;
}

extern "C" int64_t b200w_tom_shared_scale_offset(int64_t op_rows, int64_t n) {
  return round_up((int64_t)B200W_I8_SLICES * op_rows * b200w_padded_n(n), 256);
}
extern "C" int64_t b200w_tom_shared_bytes(int64_t op_rows, int64_t n) {
  return b200w_tom_shared_scale_offset(op_rows, n) + 2 * op_rows * 8;
}

int b200w_dev_intramodular(const double* dA, int64_t n, const int32_t* d_labels,
                           const uint8_t* d_small, double* d_total, double* d_within,
                           cudaStream_t st) {
  intramodular_kernel<<<(unsigned)n, 256, 0, st>>>(dA, n, d_labels, d_small, d_total, d_within);
  B200W_LAUNCHED();
  return 0;
}
