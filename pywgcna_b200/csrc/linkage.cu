// Average-linkage hierarchical clustering of the genes on the device: the step that follows the TOM in
// WGCNA.findModules, `linkage(squareform(dissTOM), method="average")` (PyWGCNA/wgcna.py:326-328), reproducing
// scipy's result exactly -- merge order, heights and cluster sizes -- so that cutreeHybrid sees the same tree.
//
// Third-party algorithm restated (scipy==1.14.0 pinned by the reference, scipy/cluster/_hierarchy.pyx):
// `nn_chain`: walk a chain of nearest neighbours until two clusters are mutual nearest neighbours, merge them,
// update the distances of the merged cluster with the Lance-Williams rule of the average method
// d(new, i) = (n_x d(x, i) + n_y d(y, i)) / (n_x + n_y), continue from the rest of the chain; then sort the merges
// stably by height and relabel them by union-find (`label`).  The tie rules that decide the tree when many
// distances are equal (round(1 - TOM, 8) has many) are kept: a scan takes the FIRST index of the strictly
// smallest distance, and the previous chain element wins a tie.  oracle/restated.py::linkage_average is the same
// restatement in numpy, checked bit for bit against scipy; tests compare this kernel with both.
//
// The walk is sequential (about 3 n scans of a distance row and n row updates), each step is a parallel
// reduction / update over n columns: one thread-block CLUSTER of 8 CTAs x 1024 threads shares every step -- each CTA
// owns a slice of the columns, candidates are exchanged through distributed shared memory, one cluster barrier per
// scan and one per merge -- instead of a kernel launch per step.  The n x n float64 dissimilarity stays in HBM
// (square form, symmetric, updated in place: row y contiguously, column y strided).
#include <cooperative_groups.h>

#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b200w {

constexpr int kLinkCtas = 8;
constexpr int kLinkThreads = 1024;

struct LinkCand {
  double d;
  long long i;
};

__device__ __forceinline__ bool cand_less(double d0, long long i0, double d1, long long i1) {
  return d0 < d1 || (d0 == d1 && i0 < i1);
}

// alive[] : one bit per cluster, replicated in every CTA's shared memory (all CTAs apply the same updates)
__global__ void __cluster_dims__(kLinkCtas, 1, 1) __launch_bounds__(kLinkThreads, 1)
linkage_average_kernel(double* __restrict__ D, const long long n, int* __restrict__ csize,
                       int* __restrict__ chain_all, double* __restrict__ Zraw) {
  extern __shared__ unsigned int alive[];  // ceil(n / 32) words
  __shared__ LinkCand slots[2][kLinkCtas];
  __shared__ LinkCand warp_best[kLinkThreads / 32];
  __shared__ long long s_x, s_y;
  __shared__ double s_cur;
  __shared__ int s_merge;  // 1: x and y are mutual nearest neighbours
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long words = (n + 31) / 32;
  for (long long w = tid; w < words; w += kLinkThreads) {
    const long long left = n - w * 32;
    alive[w] = left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
  }
  for (long long i = (long long)rank * kLinkThreads + tid; i < n; i += (long long)kLinkCtas * kLinkThreads) csize[i] = 1;
  int* chain = chain_all + (long long)rank * n;  // this CTA's replica of the chain (thread 0 only)
  // this CTA's slice of the columns
  const long long per = (n + kLinkCtas - 1) / kLinkCtas;
  const long long c0 = per * rank, c1 = (c0 + per < n) ? c0 + per : n;
  long long chain_len = 0, first_alive = 0;
  unsigned int parity = 0;
  __syncthreads();
  cluster.sync();

  for (long long k = 0; k < n - 1; ++k) {
    if (tid == 0 && chain_len == 0) {
      while (!((alive[first_alive >> 5] >> (first_alive & 31)) & 1u)) ++first_alive;
      chain[0] = (int)first_alive;
      chain_len = 1;
    }
    // ---- walk the chain until the top two elements are mutual nearest neighbours ----
    while (true) {
      if (tid == 0) {
        s_x = chain[chain_len - 1];
        s_y = chain_len > 1 ? chain[chain_len - 2] : -1;
      }
      __syncthreads();
      const long long x = s_x, yprev = s_y;
      const double* row = D + x * n;
      double bd = INFINITY;
      long long bi = n;
      for (long long i = c0 + tid; i < c1; i += kLinkThreads) {
        if (i == x || !((alive[i >> 5] >> (i & 31)) & 1u)) continue;
        const double d = __ldcg(row + i);
        if (d < bd) { bd = d; bi = i; }  // ascending i per thread: the first index of the smallest distance
      }
      for (int o = 16; o; o >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, bd, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cand_less(od, oi, bd, bi)) { bd = od; bi = oi; }
      }
      if (lane == 0) { warp_best[warp].d = bd; warp_best[warp].i = bi; }
      __syncthreads();
      if (warp == 0) {
        bd = warp_best[lane].d;
        bi = warp_best[lane].i;
        for (int o = 16; o; o >>= 1) {
          const double od = __shfl_xor_sync(0xffffffffu, bd, o);
          const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (cand_less(od, oi, bd, bi)) { bd = od; bi = oi; }
        }
        if (lane < kLinkCtas) {  // publish this CTA's candidate in every CTA's slot table
          LinkCand* remote = cluster.map_shared_rank(&slots[parity][rank], lane);
          remote->d = bd;
          remote->i = bi;
        }
      }
      cluster.sync();
      if (tid == 0) {
        double gd = slots[parity][0].d;
        long long gi = slots[parity][0].i;
        for (int c = 1; c < kLinkCtas; ++c)
          if (cand_less(slots[parity][c].d, slots[parity][c].i, gd, gi)) { gd = slots[parity][c].d; gi = slots[parity][c].i; }
        // scipy: current_min starts at the distance to the previous chain element, which therefore wins ties
        double cur = INFINITY;
        long long y = -1;
        if (yprev >= 0) { cur = __ldcg(row + yprev); y = yprev; }
        if (gd < cur) { cur = gd; y = gi; }
        if (chain_len > 1 && y == yprev) {
          s_merge = 1;
        } else {
          s_merge = 0;
          chain[chain_len] = (int)y;
          ++chain_len;
        }
        s_y = y;
        s_cur = cur;
      }
      parity ^= 1u;
      __syncthreads();
      if (s_merge) break;
    }
    // ---- merge x and y (convention of fastcluster / scipy: x < y, x is dropped, y becomes the new cluster) ----
    long long x = s_x, y = s_y;
    const double cur = s_cur;
    if (x > y) { const long long t = x; x = y; y = t; }
    const int nx = __ldcg(csize + x), ny = __ldcg(csize + y);
    if (tid == 0) {
      chain_len -= 2;
      alive[x >> 5] &= ~(1u << (x & 31));
      if (rank == 0) {
        Zraw[k * 4 + 0] = (double)x;
        Zraw[k * 4 + 1] = (double)y;
        Zraw[k * 4 + 2] = cur;
        Zraw[k * 4 + 3] = (double)(nx + ny);
      }
    }
    __syncthreads();
    {
      const double fx = (double)nx, fy = (double)ny, fs = (double)(nx + ny);
      const double* rx = D + x * n;
      double* ry = D + y * n;
      for (long long i = c0 + tid; i < c1; i += kLinkThreads) {
        if (i == y || !((alive[i >> 5] >> (i & 31)) & 1u)) continue;
        // (size_x * d_xi + size_y * d_yi) / (size_x + size_y), no fused multiply-add
        const double nd = __ddiv_rn(__dadd_rn(__dmul_rn(fx, __ldcg(rx + i)), __dmul_rn(fy, __ldcg(ry + i))), fs);
        __stcg(ry + i, nd);
        __stcg(D + i * n + y, nd);
      }
    }
    cluster.sync();  // every CTA has read csize[x], csize[y] and finished its part of the update
    if (rank == 0 && tid == 0) __stcg(csize + y, nx + ny);
    // (the next read of csize happens after at least one more cluster barrier)
  }
  cluster.sync();
}

}  // namespace b200w

using namespace b200w;

extern "C" int b200w_dev_linkage_average(double* d_D, int64_t n, double* d_Zraw, int device, void* stream) {
  if (!d_D || !d_Zraw || n < 2) return fail(B200W_E_INVALID, "linkage_average: bad args");
  if (n > 2000000) return fail(B200W_E_UNSUPPORTED, "linkage_average: n too large");
  if (int e = ensure_device(device)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  Scratch csize, chain;
  B200W_CUDA(csize.alloc((size_t)n * sizeof(int), st));
  B200W_CUDA(chain.alloc((size_t)kLinkCtas * n * sizeof(int), st));
  const size_t smem = (size_t)((n + 31) / 32) * sizeof(unsigned int);
  if (smem > 200 * 1024) return fail(B200W_E_UNSUPPORTED, "linkage_average: n too large for the alive mask");
  B200W_CUDA(cudaFuncSetAttribute(linkage_average_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linkage_average_kernel<<<kLinkCtas, kLinkThreads, smem, st>>>(d_D, (long long)n, csize.as<int>(), chain.as<int>(),
                                                                d_Zraw);
  B200W_LAUNCHED();
  return 0;
}

// scipy's post-processing of nn_chain: stable sort of the merges by height, then `label` (union-find): the two
// cluster ids of a merge become the current roots (smaller first), the merged cluster gets the next id n + i.
extern "C" int b200w_linkage_finish(double* Z, int64_t n) {
  if (!Z || n < 2) return fail(B200W_E_INVALID, "linkage_finish: bad args");
  const int64_t m = n - 1;
  std::vector<int64_t> order(m);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return Z[a * 4 + 2] < Z[b * 4 + 2]; });
  std::vector<double> S((size_t)m * 4);
  for (int64_t i = 0; i < m; ++i)
    for (int c = 0; c < 4; ++c) S[i * 4 + c] = Z[order[i] * 4 + c];
  std::vector<int64_t> parent(2 * n - 1), size(2 * n - 1, 1);
  std::iota(parent.begin(), parent.end(), 0);
  auto find = [&](int64_t a) {
    int64_t r = a;
    while (parent[r] != r) r = parent[r];
    while (parent[a] != r) { const int64_t nx = parent[a]; parent[a] = r; a = nx; }
    return r;
  };
  int64_t next = n;
  for (int64_t i = 0; i < m; ++i) {
    const int64_t a = find((int64_t)S[i * 4]), b = find((int64_t)S[i * 4 + 1]);
    Z[i * 4] = (double)(a < b ? a : b);
    Z[i * 4 + 1] = (double)(a < b ? b : a);
    Z[i * 4 + 2] = S[i * 4 + 2];
    parent[a] = parent[b] = next;
    size[next] = size[a] + size[b];
    Z[i * 4 + 3] = (double)size[next];
    ++next;
  }
  return 0;
}

// condensed (scipy squareform order) -> square symmetric with a zero diagonal, on the device
__global__ void __launch_bounds__(256)
condensed_to_square_kernel(const double* __restrict__ y, long long n, double* __restrict__ D) {
  const long long i = blockIdx.x;
  for (long long j = threadIdx.x; j < n; j += 256) {
    double v = 0.0;
    if (j > i) v = y[n * i - i * (i + 1) / 2 + (j - i - 1)];
    else if (j < i) v = y[n * j - j * (j + 1) / 2 + (i - j - 1)];
    D[i * n + j] = v;
  }
}

extern "C" int b200w_dev_condensed_to_square(const double* d_y, int64_t n, double* d_D, int device, void* stream) {
  if (!d_y || !d_D || n < 2) return fail(B200W_E_INVALID, "condensed_to_square: bad args");
  if (int e = ensure_device(device)) return e;
  condensed_to_square_kernel<<<(unsigned)n, 256, 0, (cudaStream_t)stream>>>(d_y, (long long)n, d_D);
  B200W_LAUNCHED();
  return 0;
}
