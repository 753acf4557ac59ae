// Pieces shared by the FP64 and int8 TOM kernels.
#pragma once
#include "common.cuh"

namespace b200w {

// progress of a whole-matrix symmetric TOM launch, for a host that copies finished row blocks out while
// the kernel is still running: flags[b] (host-mapped) becomes 1 when rows [b, b + 1) * rows_per_block are final
struct TomProgress {
  volatile unsigned int* flags = nullptr;
  int nb = 0;
  int64_t rows_per_block = 0;
};

// Column operand of a contraction whose adjacency is NOT symmetric (TomSimilarityFromAdj on an unchecked
// matrix, or TOMsimilarity on a matrix whose NaNs waved it through checkAdjMat): L = A @ A contracts the
// rows of A with the rows of A^T, the denominator takes the column sums k_j (wgcna.py:1145-1147).
// nan_mode != 0: no nan_to_num happened (TomSimilarityFromAdj called directly): entries whose row or
// column holds a NaN come out NaN, as numpy's A @ A and plain sums make them.
struct TomColumnOperand {
  const void* operand = nullptr;  // planes / cleaned copy of A^T, laid out like the row operand
  const double* scale = nullptr;  // int8 only
  const double* k = nullptr;      // column sums of A (NaN for a NaN column when nan_mode)
  int nan_mode = 0;
};

int tom_transpose(const double* dA, int64_t n, double* dAt, cudaStream_t st);
int tom_prepare_impl(const double* dA_rows, int64_t n, int64_t row0, int64_t nrows, int precision, void* d_operand,
                     int64_t op_rows, double* d_scale, double* d_k, int device, void* stream, int nan_mode);
int tom_compute_impl(const double* dA_rows, const void* d_operand, int64_t op_rows, const double* d_scale,
                     const double* d_k, int64_t n, int64_t row0, int64_t nrows, int tom_type, int tom_denom,
                     int out_mode, int precision, double* d_out, int device, void* stream,
                     const TomColumnOperand* col);

#ifdef __CUDACC__
// np.nan_to_num (wgcna.py:1185; default also maps +-inf to +-DBL_MAX, unreachable for a matrix that
// passed checkAdjMat) and fill_diagonal(A, 0) (wgcna.py:1143)
__device__ __forceinline__ double tom_clean(double a, int64_t i, int64_t j) {
  return (i == j || isnan(a)) ? 0.0 : a;
}

// TomSimilarityFromAdj element, wgcna.py:1148-1156, plus the 1 - TOM / round(8) hand-off of
// findModules (wgcna.py:323-324).  L = (A @ A)[i, j], a = cleaned A[i, j].
__device__ __forceinline__ double tom_value(double L, double a, double ki, double kj, bool diag,
                                            int tom_type, int tom_denom, int out_mode) {
  double t;
  if (diag) {
    t = 1.0;  // wgcna.py:1156
  } else {
    const double D = (tom_denom == B200W_DENOM_MIN) ? fmin(ki, kj) : (ki + kj) / 2.0;
    if (tom_type == B200W_TOM_UNSIGNED) t = (L + a) / (D + 1.0 - a);  // :1153
    else t = fabs(L + a) / (D + 1.0 - fabs(a));                       // :1155
  }
  if (out_mode == B200W_OUT_TOM) return t;
  t = 1.0 - t;
  if (out_mode == B200W_OUT_DISS_R8 || out_mode == B200W_OUT_DISS_R8_COND)
    t = rint(t * 1e8) / 1e8;  // numpy around(decimals=8)
  return t;
}

__host__ __device__ __forceinline__ bool out_is_condensed(int out_mode) {
  return out_mode == B200W_OUT_DISS_COND || out_mode == B200W_OUT_DISS_R8_COND;
}
// index of (i, j), i < j, in scipy's condensed distance vector
__device__ __forceinline__ int64_t condensed_index(int64_t n, int64_t i, int64_t j) {
  return n * i - i * (i + 1) / 2 + (j - i - 1);
}

__device__ __forceinline__ double block_sum_256(double v) {
  __shared__ double red_[8];
  __syncthreads();  // protect red_ across consecutive calls
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red_[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = red_[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) s += red_[w];
  return s;
}

// 1-D tile index -> (tm, tn) with row tiles taken in groups of `group`, column index fastest inside
// a group's row: concurrently resident CTAs then cover a compact block of the output and share
// operand panels in L2.
__device__ __forceinline__ void grouped_tile(int64_t pid, int64_t tiles_m, int64_t tiles_n,
                                             int group, int64_t& tm, int64_t& tn) {
  const int64_t per_group = (int64_t)group * tiles_n;
  const int64_t gid = pid / per_group;
  const int64_t first_m = gid * group;
  const int64_t gsz = (tiles_m - first_m < group) ? tiles_m - first_m : group;
  const int64_t r = pid % per_group;
  tm = first_m + r % gsz;
  tn = r / gsz;
}
#endif

}  // namespace b200w
