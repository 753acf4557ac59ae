// Correlation side of the hot path: standardise, soft-threshold sweep, adjacency.
// Reference arithmetic: np.corrcoef (numpy lib/_function_base_impl.py cov/corrcoef) as called at
// PyWGCNA/wgcna.py:882 and :1097; sweep wgcna.py:884-916; adjacency wgcna.py:1102-1111.
#include <map>
#include <mutex>
#include <vector>

#include "gemm_f64.cuh"
#include "sweep_common.cuh"

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
// K0 for GENE-MAJOR input (Xt is n x m: what DataFrame.to_numpy() hands out as an F-ordered m x n array when pandas
// keeps its block gene-major -- accepted as it is instead of a 200 MB host-side transposition).  The arithmetic is
// the same: left-to-right sum over the samples, so exactly-constant genes centre to exact zeros.  Phase 1: one thread
// per gene (its samples are contiguous; sectors are reused from L1/L2); phase 2: coalesced normalisation.
__global__ void __launch_bounds__(128)
standardize_gm_stats_kernel(const double* __restrict__ Xt, int64_t m, int64_t n, double* __restrict__ mean_out,
                            double* __restrict__ nrm_out, uint8_t* __restrict__ bad) {
  const int64_t g = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (g >= n) return;
  const double* row = Xt + g * m;
  double sum = 0.0;
  bool finite = true;
  for (int64_t s = 0; s < m; ++s) {
    const double x = row[s];
    finite = finite && isfinite(x);
    sum += x;
  }
  const double mean = sum / (double)m;
  double ss = 0.0;
  for (int64_t s = 0; s < m; ++s) {
    const double d = row[s] - mean;
    ss += d * d;
  }
  const bool is_bad = !finite || !(ss > 0.0);
  bad[g] = is_bad ? 1 : 0;
  mean_out[g] = mean;
  nrm_out[g] = is_bad ? 0.0 : sqrt(ss);
}

__global__ void __launch_bounds__(256)
standardize_gm_apply_kernel(const double* __restrict__ Xt, int64_t m, int64_t n, int64_t ldk,
                            const double* __restrict__ mean, const double* __restrict__ nrm,
                            double* __restrict__ Z) {
  const int64_t g = blockIdx.x;
  const double mu = mean[g], nr = nrm[g];
  for (int64_t s = threadIdx.x; s < ldk; s += 256) {
    double z = 0.0;
    if (s < m && nr > 0.0) z = (Xt[g * m + s] - mu) / nr;
    Z[g * ldk + s] = z;
  }
}

// -------------------------------------------------------------------------------------------
// Both kernels below use the narrow engine (gemm_f64.cuh: 128 x 64 tile, 4 warps, two CTAs per SM) and
// hand the finished tile from the accumulator registers to the epilogue through the -- by then idle --
// operand ring: each warp parks its own 64 x 32 sub-tile there and re-reads it with lane <-> column.
// That frees the 128 accumulator registers, turns the epilogue into short rolled loops (the unrolled
// register-indexed form was instruction-fetch bound in the adjacency and shuffle-latency bound in the
// sweep) and makes every global store a 256-byte row segment.
// -------------------------------------------------------------------------------------------
using Narrow = GemmCfg<2>;
constexpr int kNarrowBN = Narrow::kBNt, kNarrowThreads = Narrow::kThreads;

// park the warp's 64 x 32 accumulator sub-tile at wt[row * kPitch + col]
template <int kPitch>
__device__ __forceinline__ void park_acc(const AccF64& acc, double* wt, int lane) {
  const int g = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ch = 0; ch < 2; ++ch)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        double* p = wt + (mi * 16 + ch * 8 + g) * kPitch + ni * 8 + 2 * tig;
        if (kPitch % 2 == 0) {
          *reinterpret_cast<double2*>(p) = make_double2(acc.v[mi][ni][ch * 2], acc.v[mi][ni][ch * 2 + 1]);
        } else {
          p[0] = acc.v[mi][ni][ch * 2];
          p[1] = acc.v[mi][ni][ch * 2 + 1];
        }
      }
  __syncwarp();
}

// -------------------------------------------------------------------------------------------
// K2  fused soft-threshold sweep.  The work is the CT x RT grid of 128 x 64 correlation tiles
// (CT column tiles of the genes whose connectivity is wanted, RT row tiles of all genes), taken
// column-major and cut into one contiguous, equally long range per CTA (persistent, two CTAs per SM:
// no tail wave).  For each tile the CTA computes Z_i Z_j^T on the DMMA pipe and evaluates
// clip -> similarity transform -> diag <- 1 -> NaN skip -> the running product chain for all P powers,
// reducing over rows.  The n x n matrix is never written.
// Epilogue: a lane owns one column of the warp's sub-tile and walks its 64 rows four at a time; when every
// step of the chain is 1 or 2 (the reference's default list, and any list of consecutive or even powers)
// the P running column sums stay in registers for the whole tile -- one DMUL + one DADD per element and
// power, no shuffles; other step lists go through a rolled loop with pow() and shared-memory sums.
// Column sums of a column tile accumulate in shared memory with exactly one owner thread per
// (warp row, power, column); whenever the range crosses into the next column tile the sums are flushed to
// a partial slot, and sweep_reduce_kernel adds the slots of a column tile in CTA order: deterministic,
// no atomics.
// Symmetric mode (whole matrix, steps 1 / 2): the similarity is symmetric, so only the tiles on and above the
// diagonal are formed -- half the MMAs.  A tile strictly above the diagonal serves two sets of genes: its
// column sums go to the genes of its columns as before, and a second epilogue pass over the parked tile with
// lane <-> row gives the row sums for the genes of its rows (their column tiles never visit these rows);
// those go to a per-tile slot and sweep_reduce_sym_kernel adds, per gene, the column slots and the row
// slots in a fixed order.  The chain is evaluated twice on such a tile, so the FP64 epilogue work per
// matrix element stays what it was; the saving is the MMA half of the shared pipe.
// -------------------------------------------------------------------------------------------
constexpr int kSweepStages = 3;
// (kSweepMaxP = 20: 3 ring slots + 2 x 20 x 64 sums let two CTAs fit one SM)
constexpr int kSweepPitch = 41;  // doubles; odd, so the transposed re-read of the parked tile is 2-way at worst
#ifndef B200W_SWEEP_ROWS
#define B200W_SWEEP_ROWS 8
#endif
constexpr int kSweepRows = B200W_SWEEP_ROWS;  // rows of a column in flight per lane: independent chains

__global__ void __launch_bounds__(kNarrowThreads, 2) sweep_kernel(const SweepParams prm) {
  constexpr int kBN = kNarrowBN;
  extern __shared__ __align__(16) double smem[];
  double* colacc = smem + kSweepStages * Narrow::kSlotDoubles;  // [2][P][64]
  __shared__ uint8_t bad_col[kBN], bad_row[kBM];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  double* wt = smem + warp * 64 * kSweepPitch;  // this warp's parking space inside the ring
  const int P = prm.P;
  const int64_t n = prm.n;
  const int64_t jend = prm.col0 + prm.ncols;
  int64_t t0, t1;
  sweep_range(prm.T, gridDim.x, blockIdx.x, t0, t1);
  if (t0 >= t1) return;
  int64_t jt, it;
  if (prm.symmetric) {
    jt = sym_col_of(t0);
    it = t0 - sym_offset(jt);
  } else {
    jt = t0 / prm.RT;
    it = t0 - jt * prm.RT;
  }
  const int64_t jt_first = jt;
  int64_t cur_jt = -1;

  for (int64_t t = t0; t < t1; ++t, ++it) {
    if (it >= (prm.symmetric ? (jt >> 1) + 1 : prm.RT)) { ++jt; it = 0; }
    const int64_t j0 = prm.col0 + jt * kBN, i0 = it * kBM;
    const bool two_pass = prm.symmetric && it < (jt >> 1);  // strictly above the diagonal block
    __syncthreads();  // every warp is done with the parked tile, bad_row and (on a flush) the sums
    if (jt != cur_jt) {  // entering a new column tile: flush the previous one, reset the sums
      if (cur_jt >= 0) {
        double* dst = prm.partial + ((int64_t)blockIdx.x * prm.max_seg + (cur_jt - jt_first)) * P * kBN;
        for (int i = threadIdx.x; i < P * kBN; i += kNarrowThreads) dst[i] = colacc[i] + colacc[P * kBN + i];
        __syncthreads();
      }
      for (int i = threadIdx.x; i < 2 * P * kBN; i += kNarrowThreads) colacc[i] = 0.0;
      if (threadIdx.x < kBN) {
        int64_t j = j0 + threadIdx.x;
        bad_col[threadIdx.x] = (j < jend) ? prm.bad[j] : 1;
      }
      cur_jt = jt;
    }
    {
      int64_t i = i0 + threadIdx.x;  // 128 threads = 128 rows
      bad_row[threadIdx.x] = (i < n) ? prm.bad[i] : 1;
    }
    {
      AccF64 acc;
      acc.clear();
      // rows of the tile = genes i (summed over), columns = genes j (kept)
      gemm_f64_tile<kSweepStages, 2>(acc, smem, prm.Z, prm.Z, prm.ldk, prm.ldk, i0, n, j0, n);
      // (ends with a __syncthreads: the ring is idle, bad_row / bad_col / the zeroed sums are visible)
      park_acc<kSweepPitch>(acc, wt, lane);
    }

    const int col = wn * 32 + lane;
    const int64_t gj = j0 + col;
    const bool cbad = bad_col[col];
    double* mysum = colacc + (size_t)wm * P * kBN + col;  // [p * kBN]: this thread's own sums

    // similarity and chain start of one element: wgcna.py:884-897, NaN skipped by nansum (:915).
    // With prm.sim the similarity is also written out the way WGCNA.adjacency forms it before the power
    // (wgcna.py:1097-1107: diagonal not forced, NaN rows and columns kept), so that the adjacency of the
    // same data can be finished by an elementwise pass instead of a second correlation product.
    auto load = [&](int r, double& sv, double& c0) {
      const int row = wm * 64 + r;
      const int64_t gi = i0 + row;
      double c = wt[r * kSweepPitch + lane];
      c = fmin(1.0, fmax(-1.0, c));  // np.corrcoef clips to [-1, 1]
      sv = net_transform(c, prm.net_type);
      c0 = 1.0;
      const bool nan_pair = cbad || bad_row[row];
      if (prm.sim != nullptr) {
        const double s_out = nan_pair ? __longlong_as_double(0x7ff8000000000000LL) : sv;
        if (prm.sim_transposed) wt[r * kSweepPitch + lane] = s_out;  // stored by the transposed pass below
        else if (gi < n && gj < jend) prm.sim[gi * n + gj] = s_out;
      }
      if (gi == gj) sv = 1.0;  // wgcna.py:897 (even for a NaN gene)
      else if (nan_pair) { sv = 0.0; c0 = 0.0; }
      if (gi >= n) { sv = 0.0; c0 = 0.0; }
    };

    if (prm.simple_steps) {
      double ks[kSweepMaxP];
#pragma unroll
      for (int p = 0; p < kSweepMaxP; ++p) ks[p] = 0.0;
#pragma unroll 1
      for (int r0 = 0; r0 < 64; r0 += kSweepRows) {
        double s1[kSweepRows], s2[kSweepRows], cur[kSweepRows];
#pragma unroll
        for (int u = 0; u < kSweepRows; ++u) {
          load(r0 + u, s1[u], cur[u]);
          s2[u] = s1[u] * s1[u];
        }
#pragma unroll
        for (int p = 0; p < kSweepMaxP; ++p) {
          if (p < P) {  // wgcna.py:914  corxCur = corxPrev * corx ** step
            const bool one = prm.istep[p] == 1;
#pragma unroll
            for (int u = 0; u < kSweepRows; ++u) cur[u] *= one ? s1[u] : s2[u];
            double part[kSweepRows / 4];
#pragma unroll
            for (int q = 0; q < kSweepRows / 4; ++q)
              part[q] = (cur[4 * q] + cur[4 * q + 1]) + (cur[4 * q + 2] + cur[4 * q + 3]);
            double tot = part[0];
#pragma unroll
            for (int q = 1; q < kSweepRows / 4; ++q) tot += part[q];
            ks[p] += tot;
          }
        }
      }
#pragma unroll
      for (int p = 0; p < kSweepMaxP; ++p)
        if (p < P) mysum[p * kBN] += ks[p];
    } else {
#pragma unroll 1
      for (int r0 = 0; r0 < 64; r0 += 4) {
        double s1[4], cur[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) load(r0 + u, s1[u], cur[u]);
#pragma unroll 1
        for (int p = 0; p < P; ++p) {
          const double st = prm.step[p];
          const int ist = prm.istep[p];
#pragma unroll
          for (int u = 0; u < 4; ++u) cur[u] *= step_pow(s1[u], st, ist);
          mysum[p * kBN] += (cur[0] + cur[1]) + (cur[2] + cur[3]);
        }
      }
    }
    if (two_pass) {
      // rows lane and lane + 32 of the warp's 64, its 32 columns four at a time: row sums for the genes of
      // the tile's rows, and the mirrored half of the similarity
      double kr[2][kSweepMaxP];
#pragma unroll
      for (int p = 0; p < kSweepMaxP; ++p) kr[0][p] = kr[1][p] = 0.0;
#pragma unroll 1
      for (int c0 = 0; c0 < 32; c0 += 4) {
        double s1[8], s2[8], cur[8];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int r = h * 32 + lane, row = wm * 64 + r, colx = wn * 32 + c0 + u;
            const int64_t gi = i0 + row, gjx = j0 + colx;
            double c = wt[r * kSweepPitch + c0 + u];
            c = fmin(1.0, fmax(-1.0, c));
            double sv = net_transform(c, prm.net_type), c0v = 1.0;
            const bool nan_pair = bad_col[colx] || bad_row[row];
            const bool inside = gi < n && gjx < jend;
            if (prm.sim != nullptr && inside)
              prm.sim[gjx * n + gi] = nan_pair ? __longlong_as_double(0x7ff8000000000000LL) : sv;
            if (nan_pair || !inside) { sv = 0.0; c0v = 0.0; }
            s1[h * 4 + u] = sv;
            s2[h * 4 + u] = sv * sv;
            cur[h * 4 + u] = c0v;
          }
#pragma unroll
        for (int p = 0; p < kSweepMaxP; ++p) {
          if (p < P) {
            const bool one = prm.istep[p] == 1;
#pragma unroll
            for (int e = 0; e < 8; ++e) cur[e] *= one ? s1[e] : s2[e];
            kr[0][p] += (cur[0] + cur[1]) + (cur[2] + cur[3]);
            kr[1][p] += (cur[4] + cur[5]) + (cur[6] + cur[7]);
          }
        }
      }
      // the two warps that share rows add up through their own (now free) parking space
      __syncwarp();
#pragma unroll
      for (int p = 0; p < kSweepMaxP; ++p)
        if (p < P) {
          wt[p * 64 + lane] = kr[0][p];
          wt[p * 64 + 32 + lane] = kr[1][p];
        }
      __syncthreads();
      {
        const int row = threadIdx.x, rl = row & 63;
        const double* w0 = smem + ((row >> 6) * 2) * 64 * kSweepPitch;
        const double* w1 = w0 + 64 * kSweepPitch;
        double* dst = prm.rowpart + t * P * kBM;
        for (int p = 0; p < P; ++p) dst[p * kBM + row] = w0[p * 64 + rl] + w1[p * 64 + rl];
      }
    }
    if (prm.sim != nullptr && prm.sim_transposed) {  // lane <-> row of the sub-tile: 256-byte row segments
      __syncwarp();
#pragma unroll 1
      for (int c = 0; c < 32; ++c) {
        const int64_t gjc = j0 + wn * 32 + c;
        if (gjc >= jend) break;  // warp-uniform
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = h * 32 + lane;
          const int64_t gi = i0 + wm * 64 + r;
          if (gi < n) prm.sim[(gjc - prm.col0) * n + gi] = wt[r * kSweepPitch + c];
        }
      }
    }
  }
  __syncthreads();
  {
    double* dst = prm.partial + ((int64_t)blockIdx.x * prm.max_seg + (cur_jt - jt_first)) * P * kBN;
    for (int i = threadIdx.x; i < P * kBN; i += kNarrowThreads) dst[i] = colacc[i] + colacc[P * kBN + i];
  }
}

// k[p][j] = sum over the CTAs whose range covers column tile jt of their slot - 1   (wgcna.py:915)
__global__ void sweep_reduce_kernel(const double* __restrict__ partial, int G, int max_seg, int P,
                                    int64_t RT, int64_t T, int64_t ncols, double* __restrict__ k) {
  const int kBN = blockDim.x;  // tile columns of the sweep that wrote the slots
  const int64_t jt = blockIdx.x;
  const int col = threadIdx.x;
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

// symmetric mode: k[p][j] = column slots of column tile jt (as above) + row slots of the tiles to the right of
// gene j's 128-block, in column-tile order, - 1
__global__ void sweep_reduce_sym_kernel(const double* __restrict__ partial, const double* __restrict__ rowpart,
                                        int G, int max_seg, int P, int64_t CT, int64_t T, int64_t n,
                                        double* __restrict__ k) {
  constexpr int kBN = kNarrowBN;
  const int64_t jt = blockIdx.x;
  const int col = threadIdx.x;
  const int p = blockIdx.y;
  const int64_t j = jt * kBN + col;
  if (j >= n) return;
  const int64_t lo = sym_offset(jt), hi = sym_offset(jt + 1);
  double s = 0.0;
  for (int64_t b = lo * G / T; b < G; ++b) {
    int64_t t0, t1;
    sweep_range(T, G, b, t0, t1);
    if (t0 >= hi) break;
    if (t1 <= lo || t0 >= t1) continue;
    const int64_t seg = jt - sym_col_of(t0);
    s += partial[(((int64_t)b * max_seg + seg) * P + p) * kBN + col];
  }
  const int64_t I = j / kBM, local = j % kBM;
  for (int64_t jt2 = 2 * I + 2; jt2 < CT; ++jt2)
    s += rowpart[((sym_offset(jt2) + I) * P + p) * kBM + local];
  k[(int64_t)p * n + j] = s - 1.0;
}

// sharded symmetric schedule (b200w_dev_sweep_dist): this rank's PARTIAL connectivities of ALL genes --
//   own genes:   the column slots of their column tile (CTAs in list order) - 1        (wgcna.py:915)
//   every gene:  + the row slots of the listed tiles whose row block holds it (CSR row_off / row_idx)
// -- the caller sums the partial vectors of the ranks (all-reduce).  Fixed order, no atomics.
// Sharded symmetric sweep: this rank's partial connectivities of all genes.  The tile list comes in up to eight parts
// (launches): part q walks list[tb[q], tb[q] + Tq[q]) with Gq[q] CTAs, its column tile jt's run of the list is
// cstart[q * (CT + 1) + jt ..]; row sums by the CSR (row_off, row_idx) over list positions.
constexpr int kDistMaxParts = 8;
struct DistReduceParts {
  const double* partial[kDistMaxParts];
  int64_t tb[kDistMaxParts], T[kDistMaxParts];
  int G[kDistMaxParts], max_seg[kDistMaxParts];
  int nparts;
};
__global__ void __launch_bounds__(128)
sweep_reduce_dist_kernel(const DistReduceParts parts, const double* __restrict__ rowpart,
                         const int2* __restrict__ list, const int* __restrict__ cstart,
                         const int* __restrict__ row_off, const int* __restrict__ row_idx, int P, int64_t n,
                         int64_t col0, int64_t ncols, int CT, double* __restrict__ kpart) {
  const int64_t I = blockIdx.x;
  const int local = threadIdx.x, p = blockIdx.y;
  const int64_t j = I * 128 + local;
  if (j >= n) return;
  double s = 0.0;
  if (j >= col0 && j < col0 + ncols) {
    const int64_t jt = (j - col0) / kSweepBN;
    const int col = (int)((j - col0) % kSweepBN);
    for (int q = 0; q < parts.nparts; ++q) {
      const int64_t lo = cstart[q * (CT + 1) + jt], hi = cstart[q * (CT + 1) + jt + 1];
      const int64_t T = parts.T[q], tb = parts.tb[q];
      const int G = parts.G[q];
      if (lo >= hi || T <= 0) continue;
      for (int64_t b = (lo - tb) * G / T; b < G; ++b) {
        int64_t t0, t1;
        sweep_range(T, G, b, t0, t1);
        t0 += tb;
        t1 += tb;
        if (t0 >= hi) break;
        if (t1 <= lo || t0 >= t1) continue;
        const int64_t seg = jt - list[t0].y;
        s += parts.partial[q][(((int64_t)b * parts.max_seg[q] + seg) * P + p) * kSweepBN + col];
      }
    }
    s -= 1.0;
  }
  for (int q = row_off[I]; q < row_off[I + 1]; ++q) s += rowpart[((int64_t)row_idx[q] * P + p) * 128 + local];
  kpart[(int64_t)p * n + j] = s;
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
constexpr int kAdjStages = 3;   // narrow-engine ring slots of the adjacency (two CTAs per SM)
constexpr int kCrossStages = 4;  // wide-engine ring slots of the cross-correlation
constexpr int kAdjPitch = 33;    // odd: the transposed (mirror) reads of the parked tile stay 2-way at worst

// tiles_n counts 128-wide column tiles; a CTA takes the left or right 64 columns of one tile.
__global__ void __launch_bounds__(kNarrowThreads, 2)
adjacency_kernel(const double* __restrict__ Z, const uint8_t* __restrict__ bad, int64_t n,
                 int64_t ldk, int net_type, double power, int ipower, const double* __restrict__ d_power,
                 int64_t row0, int64_t nrows, int symmetric, int64_t tiles_n, double* __restrict__ A) {
  constexpr int kBN = kNarrowBN, kParts = 128 / kBN;
  if (d_power != nullptr) {  // the power the device-side selection left in HBM (no host round trip)
    power = *d_power;
    ipower = small_int_power(power);
  }
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int64_t tile = blockIdx.x / kParts, part = blockIdx.x % kParts;
  int64_t ti, tj;
  if (symmetric) {  // the upper triangle of 128 x 128 tiles, row by row
    int64_t t = tile, T = tiles_n;
    ti = (int64_t)floor((2.0 * T + 1.0 - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * (double)t)) / 2.0);
    while (ti > 0 && ti * T - ti * (ti - 1) / 2 > t) --ti;
    while ((ti + 1) * T - (ti + 1) * ti / 2 <= t) ++ti;
    tj = ti + (t - (ti * T - ti * (ti - 1) / 2));
  } else {
    ti = tile / tiles_n;
    tj = tile % tiles_n;
  }
  const int64_t i0 = row0 + ti * kBM, j0 = tj * 128 + part * kBN;
  const int64_t iend = row0 + nrows;
  const bool mirror = symmetric && ti != tj;
  if (j0 >= n) return;  // block-uniform

  double* wt = smem + warp * 64 * kAdjPitch;
  {
    AccF64 acc;
    acc.clear();
    gemm_f64_tile<kAdjStages, 2>(acc, smem, Z, Z, ldk, ldk, i0, n, j0, n);
    park_acc<kAdjPitch>(acc, wt, lane);
  }

  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int64_t wi0 = i0 + wm * 64, wj0 = j0 + wn * 32;
  {  // rows of the sub-tile, lane <-> column: one 256-byte store per row
    const int64_t gj = wj0 + lane;
    const bool cok = gj < n;
    const bool cbad = cok ? bad[gj] : false;
    // NaN flags of the 64 rows as two ballots (a load per row would serialise 64 global latencies)
    const unsigned rbad_lo = __ballot_sync(0xffffffffu, wi0 + lane < n && bad[wi0 + lane]);
    const unsigned rbad_hi = __ballot_sync(0xffffffffu, wi0 + 32 + lane < n && bad[wi0 + 32 + lane]);
#pragma unroll 1
    for (int r0 = 0; r0 < 64; r0 += 4) {  // four rows in flight: independent multiply chains
      if (wi0 + r0 >= iend) break;  // warp-uniform
      double a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        double c = wt[(r0 + u) * kAdjPitch + lane];
        c = fmin(1.0, fmax(-1.0, c));
        a[u] = net_transform(c, net_type);
      }
      if (ipower > 0) {
        int_pow4(a, ipower);
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) a[u] = pow(a[u], power);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int r = r0 + u;
        const int64_t gi = wi0 + r;
        const bool rbad = ((r < 32 ? rbad_lo : rbad_hi) >> (r & 31)) & 1u;
        if (cbad || rbad) a[u] = qnan;
        if (cok && gi < iend) A[(gi - row0) * n + gj] = a[u];
        if (mirror) wt[r * kAdjPitch + lane] = a[u];
      }
    }
  }
  if (mirror) {  // the transposed block, lane <-> row of the sub-tile: again 256-byte row segments
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < 32; ++c) {
      const int64_t gj = wj0 + c;
      if (gj >= n) break;  // warp-uniform
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = h * 32 + lane;
        const int64_t gi = wi0 + r;
        if (gi < n) A[gj * n + gi] = wt[r * kAdjPitch + c];
      }
    }
  }
}

// Adjacency from a similarity matrix the sweep left behind (SweepParams::sim): A <- S ** power in place, the
// same multiplication order as adjacency_kernel, so both routes give bitwise equal matrices.  HBM bound:
// 16 bytes per element.
__global__ void __launch_bounds__(256)
similarity_power_kernel(double* __restrict__ A, int64_t total, double power, int ipower,
                        const double* __restrict__ d_power) {
  if (d_power != nullptr) {
    power = *d_power;
    ipower = small_int_power(power);
  }
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t e0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; e0 < total; e0 += stride) {
    double a[4];
    if (e0 + 4 <= total) {
      const double2 lo = *reinterpret_cast<const double2*>(A + e0), hi = *reinterpret_cast<const double2*>(A + e0 + 2);
      a[0] = lo.x; a[1] = lo.y; a[2] = hi.x; a[3] = hi.y;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = (e0 + u < total) ? A[e0 + u] : 0.0;
    }
    bool nan[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) nan[u] = isnan(a[u]);
    if (ipower > 0) {
      int_pow4(a, ipower);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = pow(a[u], power);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (nan[u]) a[u] = qnan;
    if (e0 + 4 <= total) {
      *reinterpret_cast<double2*>(A + e0) = make_double2(a[0], a[1]);
      *reinterpret_cast<double2*>(A + e0 + 2) = make_double2(a[2], a[3]);
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e0 + u < total) A[e0 + u] = a[u];
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
  gemm_f64_tile<kCrossStages>(acc, smem, Zx, Zy, ldk, ldk, i0, n, j0, M);
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
  const size_t smem = gemm_smem_bytes(kCrossStages);
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

extern "C" int b200w_dev_standardize_gm(const double* dXt, int64_t m, int64_t n, double* dZ, uint8_t* d_bad,
                                        int device, void* stream) {
  if (!dXt || !dZ || !d_bad || m < 1 || n < 1) return fail(B200W_E_INVALID, "standardize_gm: bad args");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  Scratch mean, nrm;
  B200W_CUDA(mean.alloc((size_t)n * 8, st));
  B200W_CUDA(nrm.alloc((size_t)n * 8, st));
  standardize_gm_stats_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(dXt, m, n, mean.as<double>(),
                                                                          nrm.as<double>(), d_bad);
  B200W_LAUNCHED();
  standardize_gm_apply_kernel<<<(unsigned)n, 256, 0, st>>>(dXt, m, n, b200w_padded_k(m), mean.as<double>(),
                                                           nrm.as<double>(), dZ);
  B200W_LAUNCHED();
  return 0;
}

static int sweep_impl(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, const double* h_steps, int P,
                      int net_type, int64_t col0, int64_t ncols, double* d_k, double* d_sim, int device,
                      void* stream);

extern "C" int b200w_dev_sweep(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                               const double* h_steps, int P, int net_type, int64_t col0,
                               int64_t ncols, double* d_k, int device, void* stream) {
  return sweep_impl(dZ, d_bad, m, n, h_steps, P, net_type, col0, ncols, d_k, nullptr, device, stream);
}

extern "C" int b200w_dev_sweep_similarity(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                                          const double* h_steps, int P, int net_type, int64_t col0,
                                          int64_t ncols, double* d_k, double* d_sim, int device, void* stream) {
  if (!d_sim) return fail(B200W_E_INVALID, "sweep_similarity: bad args");
  return sweep_impl(dZ, d_bad, m, n, h_steps, P, net_type, col0, ncols, d_k, d_sim, device, stream);
}

static int adjacency_from_similarity_impl(double* d_sim_inout, int64_t nrows, int64_t n, double power,
                                          const double* d_power, int device, void* stream);

extern "C" int b200w_dev_adjacency_from_similarity(double* d_sim_inout, int64_t nrows, int64_t n, double power,
                                                   int device, void* stream) {
  if (!(power > 0.0)) return fail(B200W_E_INVALID, "adjacency_from_similarity: bad args");
  return adjacency_from_similarity_impl(d_sim_inout, nrows, n, power, nullptr, device, stream);
}

extern "C" int b200w_dev_adjacency_from_similarity_dp(double* d_sim_inout, int64_t nrows, int64_t n,
                                                      const double* d_power, int device, void* stream) {
  if (!d_power) return fail(B200W_E_INVALID, "adjacency_from_similarity_dp: bad args");
  return adjacency_from_similarity_impl(d_sim_inout, nrows, n, 1.0, d_power, device, stream);
}

static int adjacency_from_similarity_impl(double* d_sim_inout, int64_t nrows, int64_t n, double power,
                                          const double* d_power, int device, void* stream) {
  if (!d_sim_inout || n < 1 || nrows < 1 || nrows > n)
    return fail(B200W_E_INVALID, "adjacency_from_similarity: bad args");
  if (int e = ensure_device(device)) return e;
  int sms = 0;
  B200W_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int ipower = small_int_power(power);
  const int64_t total = nrows * n, per_block = 256 * 4;
  int64_t blocks = (total + per_block - 1) / per_block;
  if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
  similarity_power_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_sim_inout, total, power, ipower,
                                                                              d_power);
  B200W_LAUNCHED();
  return 0;
}

static int sweep_impl(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, const double* h_steps, int P,
                      int net_type, int64_t col0, int64_t ncols, double* d_k, double* d_sim, int device,
                      void* stream) {
  if (!dZ || !d_bad || !h_steps || !d_k || P < 1 || n < 1 || col0 < 0 || ncols < 1 ||
      col0 + ncols > n)
    return fail(B200W_E_INVALID, "sweep: bad args");
  if (net_type < 0 || net_type > 2) return fail(B200W_E_INVALID, "sweep: bad network type");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldk = b200w_padded_k(m);
  constexpr int kBN = kNarrowBN;
  const int64_t CT = (ncols + kBN - 1) / kBN, RT = (n + kBM - 1) / kBM;
  int sms = 0;
  B200W_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  static const bool sym_allowed = [] {
    const char* e = getenv("B200W_SWEEP_SYMMETRIC");
    return !(e && e[0] == '0');
  }();

  const size_t smem_max = Narrow::smem_bytes(kSweepStages) + (size_t)2 * kSweepMaxP * kBN * 8;
  B200W_CUDA(cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_max));
  for (int p0 = 0; p0 < P; p0 += kSweepMaxP) {
    // the chain restarts from 1 for each chunk of 20 powers, so the first step of a later chunk is
    // the whole power reached so far (pow() with a large exponent: the value the reference's chain
    // approximates to ~1e-15).
    const int Pc = (P - p0 < kSweepMaxP) ? P - p0 : kSweepMaxP;
    SweepParams prm = {};
    prm.Z = dZ; prm.bad = d_bad; prm.n = n; prm.ldk = ldk; prm.col0 = col0; prm.ncols = ncols;
    prm.RT = RT; prm.P = Pc; prm.net_type = net_type;
    prm.sim = (p0 == 0) ? d_sim : nullptr;  // one launch writes it
    prm.sim_transposed = (col0 == 0 && ncols == n) ? 0 : 1;
    double base = 0.0;
    for (int q = 0; q < p0; ++q) base += h_steps[q];
    prm.simple_steps = 1;
    for (int q = 0; q < kSweepMaxP; ++q) { prm.step[q] = 1.0; prm.istep[q] = 1; }
    for (int q = 0; q < Pc; ++q) {
      double s = h_steps[p0 + q] + (q == 0 ? base : 0.0);
      prm.step[q] = s;
      prm.istep[q] = (s == (double)(int)s && s >= 1.0 && s <= 4.0) ? (int)s : 0;
      if (prm.istep[q] != 1 && prm.istep[q] != 2) prm.simple_steps = 0;
    }
    // the tcgen05 engine (corr_i8.cu) takes every plain chain; other step lists stay on the DMMA engine
    const bool use_i8 = prm.simple_steps && corr_i8_enabled(m, n);
    const int64_t resident = (use_i8 ? 1 : 2) * (int64_t)sms;  // persistent CTAs per launch
    prm.adj_power = 1.0; prm.adj_ipower = 1; prm.d_adj_power = nullptr; prm.zscale = nullptr;
    // whole matrix and a plain chain: only the tiles on and above the diagonal
    bool symmetric = sym_allowed && col0 == 0 && ncols == n && prm.simple_steps;
    if (symmetric) {  // the per-tile row-sum slots must stay a modest part of the device memory
      size_t free_b = 0, total_b = 0;
      B200W_CUDA(cudaMemGetInfo(&free_b, &total_b));
      if ((size_t)sym_offset(CT) * Pc * kBM * sizeof(double) > total_b / 8) symmetric = false;
    }
    const int64_t T = symmetric ? sym_offset(CT) : CT * RT;
    const int G = (int)(T < resident ? T : resident);  // persistent: two CTAs per SM, equal tile ranges
    int max_seg = 1;  // column tiles one CTA range can touch
    for (int b = 0; b < G; ++b) {
      int64_t t0, t1;
      sweep_range(T, G, b, t0, t1);
      if (t0 >= t1) continue;
      const int64_t first = symmetric ? sym_col_of(t0) : t0 / RT, last = symmetric ? sym_col_of(t1 - 1) : (t1 - 1) / RT;
      if (last - first + 1 > max_seg) max_seg = (int)(last - first + 1);
    }
    prm.T = T; prm.max_seg = max_seg; prm.symmetric = symmetric ? 1 : 0;
    Scratch part, rowpart;
    B200W_CUDA(part.alloc((size_t)G * max_seg * Pc * kBN * sizeof(double), st));
    prm.partial = part.as<double>();
    prm.rowpart = nullptr;
    if (symmetric) {
      B200W_CUDA(rowpart.alloc((size_t)T * Pc * kBM * sizeof(double), st));
      prm.rowpart = rowpart.as<double>();
    }
    if (use_i8) {
      if (int e = sweep_i8_launch(prm, m, CT, Pc, G, device, st)) return e;
    } else {
      const size_t smem = Narrow::smem_bytes(kSweepStages) + (size_t)2 * Pc * kBN * 8;
      sweep_kernel<<<G, kNarrowThreads, smem, st>>>(prm);
      B200W_LAUNCHED();
    }
    dim3 rgrid((unsigned)CT, (unsigned)Pc);
    if (symmetric)
      sweep_reduce_sym_kernel<<<rgrid, kBN, 0, st>>>(prm.partial, prm.rowpart, G, max_seg, Pc, CT, T, n,
                                                     d_k + (int64_t)p0 * ncols);
    else
      sweep_reduce_kernel<<<rgrid, kBN, 0, st>>>(prm.partial, G, max_seg, Pc, RT, T, ncols,
                                                 d_k + (int64_t)p0 * ncols);
    B200W_LAUNCHED();
  }
  return 0;
}

static int adjacency_impl(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, int net_type, double power,
                          const double* d_power, int64_t row0, int64_t nrows, double* dA, int device, void* stream);

extern "C" int b200w_dev_adjacency(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                                   int net_type, double power, int64_t row0, int64_t nrows,
                                   double* dA, int device, void* stream) {
  return adjacency_impl(dZ, d_bad, m, n, net_type, power, nullptr, row0, nrows, dA, device, stream);
}

extern "C" int b200w_dev_adjacency_dp(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n,
                                      int net_type, const double* d_power, int64_t row0, int64_t nrows,
                                      double* dA, int device, void* stream) {
  if (!d_power) return fail(B200W_E_INVALID, "adjacency_dp: bad args");
  return adjacency_impl(dZ, d_bad, m, n, net_type, 1.0, d_power, row0, nrows, dA, device, stream);
}

static int adjacency_impl(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, int net_type, double power,
                          const double* d_power, int64_t row0, int64_t nrows, double* dA, int device, void* stream) {
  if (!dZ || !d_bad || !dA || n < 1 || row0 < 0 || nrows < 1 || row0 + nrows > n)
    return fail(B200W_E_INVALID, "adjacency: bad args");
  if (net_type < 0 || net_type > 2) return fail(B200W_E_INVALID, "adjacency: bad network type");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (corr_i8_enabled(m, n)) {
    // tcgen05 engine: the sweep's tile walk without a chain (P = 0), the output written as s ** power
    int sms = 0;
    B200W_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    SweepParams prm = {};
    prm.Z = dZ; prm.bad = d_bad; prm.n = n; prm.ldk = b200w_padded_k(m); prm.col0 = row0; prm.ncols = nrows;
    prm.RT = (n + kBM - 1) / kBM; prm.P = 0; prm.net_type = net_type; prm.simple_steps = 1;
    prm.sim = dA; prm.symmetric = (row0 == 0 && nrows == n) ? 1 : 0; prm.sim_transposed = prm.symmetric ? 0 : 1;
    prm.adj_power = power; prm.adj_ipower = small_int_power(power); prm.d_adj_power = d_power;
    const int64_t CT = (nrows + kSweepBN - 1) / kSweepBN;
    prm.T = prm.symmetric ? sym_offset(CT) : CT * prm.RT;
    prm.max_seg = 1;
    const int G = (int)(prm.T < sms ? prm.T : sms);
    return sweep_i8_launch(prm, m, CT, 0, G, device, st);
  }
  const size_t smem = Narrow::smem_bytes(kAdjStages);
  B200W_CUDA(cudaFuncSetAttribute(adjacency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
  int ipower = small_int_power(power);
  const int64_t row_tiles = (nrows + kBM - 1) / kBM, col_tiles = (n + 127) / 128;
  const int symmetric = (row0 == 0 && nrows == n) ? 1 : 0;
  const int64_t blocks = 2 * (symmetric ? col_tiles * (col_tiles + 1) / 2 : row_tiles * col_tiles);
  if (blocks > 2147483647LL) return fail(B200W_E_UNSUPPORTED, "adjacency: too many tiles");
  adjacency_kernel<<<(unsigned)blocks, kNarrowThreads, smem, st>>>(dZ, d_bad, n, b200w_padded_k(m),
                                                                  net_type, power, ipower, d_power, row0,
                                                                  nrows, symmetric, col_tiles, dA);
  B200W_LAUNCHED();
  return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// Sharded SYMMETRIC sweep (multi-GPU): the unordered pairs of 128-gene blocks are dealt out over the ranks and the rank
// that gets a pair computes it ONCE, as the tile (rows = the other block, columns = its own block): column sums for its
// own genes, row sums for the genes of the other block, the similarity transposed into its own rows -- and in the other
// orientation to the owner of the rows.  Half the MMAs and half the power chains of the column-block schedule; the
// ranks' partial connectivities are summed by the caller.  Two ways to reach the row owner:
//   direct (two ranks): pair {a, b} goes to the owner of a if a + b is even, else to the owner of b; the kernel stores
//     into the peer's block over NVLink (measured: free with ONE peer, but 4.5x the kernel's time with three);
//   staged (more ranks): between ranks r < q, r takes the pairs whose q-block lies in the lower half of q's blocks and
//     q the others, so that what a rank owes a peer is ONE rectangle; the kernel writes it into a local staging
//     rectangle, the tiles bound for peers run first, and the copy engines push the rectangles (735 GB/s per pair)
//     while a second launch does the rank's own-own pairs.
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct DistRegion {  // what this rank computes for the rows of peer o: global rows [r0, r1) x global columns [c0, c1)
  int64_t r0 = 0, r1 = 0, c0 = 0, c1 = 0;
};
struct DistPlan {
  int2* d_list = nullptr;
  int *d_cstart = nullptr, *d_row_off = nullptr, *d_row_idx = nullptr;
  std::vector<int2> list;
  int nparts = 1;
  int64_t T[kDistMaxParts] = {0};  // staged: part d - 1 = rows of rank + d (d = 1 .. world - 1), last part = own rows
  DistRegion region[8];
  cudaStream_t push_stream = nullptr;
  cudaEvent_t ev_part[kDistMaxParts] = {nullptr}, ev_pushed = nullptr;  // ev_part[q]: part q has been computed
  bool ready = false;
};
std::map<std::vector<int64_t>, DistPlan> g_dist_plans;
std::mutex g_dist_mu;

inline int64_t dist_ncols(int64_t n, int rank, int64_t per) {
  const int64_t col0 = (int64_t)rank * per;
  return col0 >= n ? 0 : (col0 + per <= n ? per : n - col0);
}

// Does `rank` form the tile (rows = block a, columns inside block gb of its own column block)?
inline bool dist_mine(int64_t a, int64_t gb, int rank, int64_t per, int64_t n, bool staged) {
  const int64_t oa = (a * 128) / per;
  if (a == gb) return true;
  if (oa == rank) return a < gb;  // a pair of two own blocks: once, in the upper orientation
  if (!staged) {
    const int64_t lo = a < gb ? a : gb, hi = a < gb ? gb : a;
    return (((lo + hi) & 1) == 0 ? (lo * 128) / per : (hi * 128) / per) == rank;
  }
  // staged: the split runs through the blocks of the HIGHER rank of the two
  const int hi_rank = oa > rank ? (int)oa : rank;
  const int64_t b0 = (int64_t)hi_rank * per / 128;
  const int64_t b1 = b0 + (dist_ncols(n, hi_rank, per) + 127) / 128;
  const int64_t mid = b0 + (b1 - b0 + 1) / 2;
  return oa > rank ? (a < mid) : (gb >= mid);
}

// The tiles of one rank; x = 128-row block, y = 64-column tile of the rank's column block; column-major inside a part,
// cstart[q * (CT + 1) + jt] = start of column tile jt in part q.  Staged mode: one part (launch) per peer in the order
// rank + 1, rank + 2, ... -- its rectangle is pushed while the next part runs -- and the own rows last; direct mode:
// everything in part 0.  Returns the number of parts.
int dist_tile_list(int64_t n, int world, int rank, int64_t per, bool staged, std::vector<int2>& list,
                   std::vector<int>& cstart, int64_t* T) {
  const int64_t col0 = (int64_t)rank * per, ncols = dist_ncols(n, rank, per);
  const int64_t B = (n + 127) / 128, CT = (ncols + kSweepBN - 1) / kSweepBN;
  const int nparts = staged ? world : 1;
  cstart.assign((size_t)nparts * (CT + 1), 0);
  list.clear();
  for (int part = 0; part < nparts; ++part) {
    const int64_t target = staged ? (rank + part + 1) % world : -1;  // (the last part: rank itself)
    const size_t before = list.size();
    for (int64_t jt = 0; jt < CT; ++jt) {
      cstart[part * (CT + 1) + jt] = (int)list.size();
      const int64_t gb = (col0 + jt * kSweepBN) / 128;
      for (int64_t a = 0; a < B; ++a) {
        if (staged && (a * 128) / per != target) continue;
        if (dist_mine(a, gb, rank, per, n, staged)) list.push_back(make_int2((int)a, (int)jt));
      }
    }
    cstart[part * (CT + 1) + CT] = (int)list.size();
    T[part] = (int64_t)(list.size() - before);
  }
  return nparts;
}

// The rectangle that `rank` owes peer o in staged mode (empty in direct mode, for o == rank and for ranks without genes)
DistRegion dist_region(int64_t n, int world, int rank, int64_t per, int o, bool staged) {
  DistRegion rg;
  (void)world;
  const int64_t col0 = (int64_t)rank * per, ncols = dist_ncols(n, rank, per);
  if (o == rank || !staged) return rg;
  const int64_t o_rows = dist_ncols(n, o, per);
  if (o_rows == 0 || ncols == 0) return rg;
  const int hi_rank = o > rank ? o : rank;
  const int64_t b0 = (int64_t)hi_rank * per / 128, b1 = b0 + (dist_ncols(n, hi_rank, per) + 127) / 128;
  const int64_t mid = b0 + (b1 - b0 + 1) / 2;
  if (o > rank) {  // rows: the lower half of o's blocks; columns: all of mine
    rg.r0 = (int64_t)o * per;
    rg.r1 = mid * 128 < n ? mid * 128 : n;
    rg.c0 = col0;
    rg.c1 = col0 + ncols;
  } else {  // rows: all of o's; columns: the upper half of my blocks
    rg.r0 = (int64_t)o * per;
    rg.r1 = rg.r0 + o_rows;
    rg.c0 = mid * 128;
    rg.c1 = col0 + ncols;
  }
  if (rg.r1 <= rg.r0 || rg.c1 <= rg.c0) rg = DistRegion();
  return rg;
}

int dist_max_seg(const std::vector<int2>& list, int64_t tb, int64_t T, int G) {
  int max_seg = 1;
  for (int b = 0; b < G; ++b) {
    int64_t t0, t1;
    sweep_range(T, G, b, t0, t1);
    if (t0 >= t1) continue;
    const int seg = list[tb + t1 - 1].y - list[tb + t0].y + 1;
    if (seg > max_seg) max_seg = seg;
  }
  return max_seg;
}

bool dist_staged(int world) {
  static const char* e = getenv("B200WGCNA_SWEEP_PUSH");  // development: "direct" / "staged" for any number of ranks
  if (e && e[0] == 'd') return false;
  if (e && e[0] == 's') return true;
  return world > 2;
}
}  // namespace

// Host only (no device needed): the tile list of b200w_dev_sweep_dist for one rank, as (row block, column tile) pairs.
// Returns the number of tiles; writes them when capacity suffices.  For tests of the schedule.
extern "C" int64_t b200w_sweep_dist_tiles(int64_t n, int world, int rank, int64_t per, int32_t* out_xy,
                                          int64_t capacity) {
  if (n < 1 || world < 2 || world > 8 || rank < 0 || rank >= world || per < 256 || per % 256) return -1;
  std::vector<int2> list;
  std::vector<int> cstart;
  int64_t T[kDistMaxParts];
  dist_tile_list(n, world, rank, per, dist_staged(world), list, cstart, T);
  if (out_xy != nullptr && capacity >= (int64_t)list.size())
    for (size_t t = 0; t < list.size(); ++t) {
      out_xy[2 * t] = list[t].x;
      out_xy[2 * t + 1] = list[t].y;
    }
  return (int64_t)list.size();
}

// Host only: the rectangles of the staged mode -- out[4 * o .. ] = {row0, row1, col0, col1} (global gene indices) of what
// `rank` computes for peer o and the copy engines push into o's block; all zero where there is none.  Returns 1 in staged
// mode, 0 in direct mode (the kernel stores into the peers itself), -1 for bad arguments.
extern "C" int b200w_sweep_dist_regions(int64_t n, int world, int rank, int64_t per, int64_t* out) {
  if (n < 1 || world < 2 || world > 8 || rank < 0 || rank >= world || per < 256 || per % 256 || !out) return -1;
  const bool staged = dist_staged(world);
  for (int o = 0; o < world; ++o) {
    const DistRegion rg = dist_region(n, world, rank, per, o, staged);
    out[4 * o] = rg.r0; out[4 * o + 1] = rg.r1; out[4 * o + 2] = rg.c0; out[4 * o + 3] = rg.c1;
  }
  return staged ? 1 : 0;
}

extern "C" int b200w_dev_sweep_dist(const double* dZ, const uint8_t* d_bad, int64_t m, int64_t n, const double* h_steps,
                                    int P, int net_type, int world, int rank, int64_t per, void* const* peer_sim,
                                    double* d_kpart, int device, void* stream) {
  if (!dZ || !d_bad || !h_steps || !d_kpart || !peer_sim || P < 1 || n < 1 || world < 2 || world > 8 || rank < 0 ||
      rank >= world || per < 256 || per % 256)
    return fail(B200W_E_INVALID, "sweep_dist: bad args");
  if (net_type < 0 || net_type > 2) return fail(B200W_E_INVALID, "sweep_dist: bad network type");
  if (P > kSweepMaxP) return fail(B200W_E_UNSUPPORTED, "sweep_dist: at most %d powers", kSweepMaxP);
  if (!corr_i8_enabled(m, n)) return fail(B200W_E_UNSUPPORTED, "sweep_dist: needs the tcgen05 correlation engine");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  SweepParams prm = {};
  for (int q = 0; q < kSweepMaxP; ++q) { prm.step[q] = 1.0; prm.istep[q] = 1; }
  for (int q = 0; q < P; ++q) {
    const double s = h_steps[q];
    prm.step[q] = s;
    prm.istep[q] = (s == 1.0) ? 1 : (s == 2.0 ? 2 : 0);
    if (!prm.istep[q]) return fail(B200W_E_UNSUPPORTED, "sweep_dist: power steps must be 1 or 2");
  }
  for (int r = 0; r < world; ++r)
    if (!peer_sim[r]) return fail(B200W_E_INVALID, "sweep_dist: peer_sim[%d] is NULL", r);
  const bool staged = dist_staged(world);
  const int64_t col0 = (int64_t)rank * per, ncols = dist_ncols(n, rank, per);
  const int64_t B = (n + 127) / 128, CT = (ncols + kSweepBN - 1) / kSweepBN;
  int sms = 0;
  B200W_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  std::lock_guard<std::mutex> lk(g_dist_mu);
  DistPlan& plan = g_dist_plans[{(int64_t)device, n, (int64_t)world, (int64_t)rank, per, (int64_t)staged}];
  if (!plan.ready) {
    std::vector<int> cstart;
    plan.nparts = dist_tile_list(n, world, rank, per, staged, plan.list, cstart, plan.T);
    const int64_t Tall = (int64_t)plan.list.size();
    std::vector<int> row_off(B + 1, 0), row_idx;
    {  // two-pass tiles (off the diagonal block) by row block, in list order
      std::vector<std::vector<int>> by_row(B);
      for (int64_t t = 0; t < Tall; ++t) {
        const int64_t a = plan.list[t].x;
        if (a != (col0 + (int64_t)plan.list[t].y * kSweepBN) / 128) by_row[a].push_back((int)t);
      }
      for (int64_t a = 0; a < B; ++a) {
        row_off[a] = (int)row_idx.size();
        row_idx.insert(row_idx.end(), by_row[a].begin(), by_row[a].end());
      }
      row_off[B] = (int)row_idx.size();
    }
    for (int o = 0; o < world; ++o) plan.region[o] = dist_region(n, world, rank, per, o, staged);
    B200W_CUDA(cudaMalloc(&plan.d_list, (plan.list.size() + 1) * sizeof(int2)));
    B200W_CUDA(cudaMalloc(&plan.d_cstart, cstart.size() * sizeof(int)));
    B200W_CUDA(cudaMalloc(&plan.d_row_off, row_off.size() * sizeof(int)));
    B200W_CUDA(cudaMalloc(&plan.d_row_idx, (row_idx.size() + 1) * sizeof(int)));
    B200W_CUDA(cudaMemcpy(plan.d_list, plan.list.data(), plan.list.size() * sizeof(int2), cudaMemcpyHostToDevice));
    B200W_CUDA(cudaMemcpy(plan.d_cstart, cstart.data(), cstart.size() * sizeof(int), cudaMemcpyHostToDevice));
    B200W_CUDA(cudaMemcpy(plan.d_row_off, row_off.data(), row_off.size() * sizeof(int), cudaMemcpyHostToDevice));
    B200W_CUDA(cudaMemcpy(plan.d_row_idx, row_idx.data(), row_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (staged) {
      B200W_CUDA(cudaStreamCreateWithFlags(&plan.push_stream, cudaStreamNonBlocking));
      for (int q = 0; q < plan.nparts; ++q)
        B200W_CUDA(cudaEventCreateWithFlags(&plan.ev_part[q], cudaEventDisableTiming));
      B200W_CUDA(cudaEventCreateWithFlags(&plan.ev_pushed, cudaEventDisableTiming));
    }
    plan.ready = true;
  }
  const int64_t Tall = (int64_t)plan.list.size();
  if (Tall == 0) {
    B200W_CUDA(cudaMemsetAsync(d_kpart, 0, (size_t)P * n * sizeof(double), st));
    return 0;
  }
  // direct mode: one launch over the whole list; staged: one launch per peer, then the own rows
  const int nparts = plan.nparts;
  DistReduceParts parts = {};
  parts.nparts = nparts;
  Scratch part[kDistMaxParts], rowpart, stage;
  int64_t tb = 0;
  for (int q = 0; q < nparts; ++q) {
    const int64_t Tq = plan.T[q];
    const int G = (int)(Tq < sms ? (Tq > 0 ? Tq : 1) : sms);
    const int max_seg = Tq > 0 ? dist_max_seg(plan.list, tb, Tq, G) : 1;
    B200W_CUDA(part[q].alloc((size_t)G * max_seg * P * kSweepBN * sizeof(double), st));
    parts.partial[q] = part[q].as<double>();
    parts.tb[q] = tb;
    parts.T[q] = Tq;
    parts.G[q] = G;
    parts.max_seg[q] = max_seg;
    tb += Tq;
  }
  B200W_CUDA(rowpart.alloc((size_t)Tall * P * 128 * sizeof(double), st));
  prm.Z = dZ; prm.bad = d_bad; prm.n = n; prm.ldk = b200w_padded_k(m); prm.col0 = col0; prm.ncols = ncols;
  prm.RT = B; prm.P = P; prm.net_type = net_type; prm.simple_steps = 1;
  prm.symmetric = 0; prm.sim_transposed = 1; prm.dist = 1; prm.per_rows = per; prm.tile_list = plan.d_list;
  prm.adj_power = 1.0; prm.adj_ipower = 1;
  prm.sim = (double*)peer_sim[rank];
  prm.rowpart = rowpart.as<double>();
  size_t stage_elems = 0;
  size_t stage_off[8] = {0};
  for (int o = 0; o < world; ++o) {
    const DistRegion& rg = plan.region[o];
    stage_off[o] = stage_elems;
    if (staged && o != rank) stage_elems += (size_t)(rg.r1 - rg.r0) * (size_t)(rg.c1 - rg.c0);
  }
  if (staged) B200W_CUDA(stage.alloc(stage_elems * sizeof(double), st));
  for (int o = 0; o < world; ++o) {
    const DistRegion& rg = plan.region[o];
    if (staged && o != rank) {
      prm.peer_sim[o] = stage.as<double>() + stage_off[o];
      prm.peer_pitch[o] = rg.c1 - rg.c0;
      prm.peer_row0[o] = rg.r0;
      prm.peer_col0[o] = rg.c0;
    } else {
      prm.peer_sim[o] = (double*)peer_sim[o];
      prm.peer_pitch[o] = n;
      prm.peer_row0[o] = (int64_t)o * per;
      prm.peer_col0[o] = 0;
    }
  }
  prm.t_begin = 0;
  prm.T = parts.T[0];
  prm.max_seg = parts.max_seg[0];
  prm.partial = part[0].as<double>();
  SweepRange more[kDistMaxParts];
  for (int q = 1; q < nparts; ++q)
    more[q - 1] = {parts.tb[q], parts.T[q], parts.G[q], parts.max_seg[q], part[q].as<double>(),
                   staged ? plan.ev_part[q - 1] : nullptr};
  if (int e = sweep_i8_launch(prm, m, CT, P, parts.G[0], device, st, more, nparts - 1)) return e;
  if (staged) {
    // a peer's rectangle is complete when its part is: the copy engines push it into the peer's block while the
    // following parts run (the last push during the own rows)
    for (int d = 1; d < world; ++d) {
      const int o = (rank + d) % world;
      const DistRegion& rg = plan.region[o];
      if (rg.r1 <= rg.r0) continue;
      B200W_CUDA(cudaStreamWaitEvent(plan.push_stream, plan.ev_part[d - 1], 0));
      double* dst = (double*)peer_sim[o] + (rg.r0 - (int64_t)o * per) * n + rg.c0;
      B200W_CUDA(cudaMemcpy2DAsync(dst, (size_t)n * sizeof(double), prm.peer_sim[o],
                                   (size_t)(rg.c1 - rg.c0) * sizeof(double), (size_t)(rg.c1 - rg.c0) * sizeof(double),
                                   (size_t)(rg.r1 - rg.r0), cudaMemcpyDeviceToDevice, plan.push_stream));
    }
    B200W_CUDA(cudaEventRecord(plan.ev_pushed, plan.push_stream));
  }
  dim3 rgrid((unsigned)B, (unsigned)P);
  sweep_reduce_dist_kernel<<<rgrid, 128, 0, st>>>(parts, rowpart.as<double>(), plan.d_list, plan.d_cstart,
                                                  plan.d_row_off, plan.d_row_idx, P, n, col0, ncols, (int)CT, d_kpart);
  B200W_LAUNCHED();
  if (staged) B200W_CUDA(cudaStreamWaitEvent(st, plan.ev_pushed, 0));  // (also before the staging area is freed)
  return 0;
}
