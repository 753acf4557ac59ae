// Pieces shared by the two correlation engines of the soft-threshold sweep / adjacency:
// corr.cu (FP64 DMMA, mma.sync) and corr_i8.cu (tcgen05 kind::i8 on digit planes of Z, TMA fed).
#pragma once
#include "common.cuh"

namespace b200w {

constexpr int kSweepMaxP = 20;  // powers per launch (column sums live in registers / one smem slot each)
constexpr int kSweepBM = 128, kSweepBN = 64;  // correlation tile of both engines

struct SweepParams {
  const double* Z;
  const uint8_t* bad;
  int64_t n, ldk;
  int64_t col0, ncols;  // genes whose connectivity this launch produces
  int64_t RT, T;        // row tiles, total tiles (CT * RT)
  int P;
  int net_type;
  int max_seg;          // partial slots per CTA
  int simple_steps;     // every istep is 1 or 2
  double* partial;      // [gridDim.x][max_seg][P][64]
  double* sim;          // optional output: the transformed similarity as adjacency() sees it (below) ...
  int sim_transposed;   // ... 0: sim[i * n + j] (whole matrix), 1: sim[(j - col0) * n + i] (this column block
                        //     as the row block of the symmetric matrix: what a rank of the sharded path keeps)
  int symmetric;        // whole matrix, only tiles on and above the diagonal (see sweep_kernel)
  double* rowpart;      // symmetric: [tile][P][128] row sums of the tiles above the diagonal
  double step[kSweepMaxP];
  int istep[kSweepMaxP];
  // tcgen05 engine only: adjacency mode (P == 0): sim receives s ** power (wgcna.py:1111) instead of s
  double adj_power;
  int adj_ipower;
  const double* d_adj_power;  // optional: the power in device memory (device-side selection)
  const double* zscale;       // per gene 2^(e - 46): Z = digits * zscale
  long long* debug;           // optional (B200W_CORR_DEBUG=1): [CTA][8] cycles per phase of epilogue warp 0 / the MMA warp
  // sharded symmetric schedule (tcgen05 engine, b200w_dev_sweep_dist): the tiles of this rank as an explicit list
  // (x = 128-row block, y = 64-column tile of the rank's column block), column-major; every tile off the diagonal block
  // serves two gene sets -- column sums for the rank's own genes, row sums for the genes of its rows (any rank's) --
  // and its similarity goes to BOTH owners: transposed into the rank's own rows, directly into the row owner's block
  // (peer_sim[owner], peer-mapped HBM) from inside the kernel
  const int2* tile_list;
  int64_t t_begin;       // this launch walks tile_list[t_begin, t_begin + T)
  double* peer_sim[8];   // where the tile's direct orientation goes, by owner of its rows: the owner's block itself
  int64_t peer_pitch[8]; // (own HBM / a peer's over NVLink: pitch n, row0 = owner * per_rows, col0 = 0) or a local staging
  int64_t peer_row0[8];  // rectangle that the copy engines push to the owner afterwards:
  int64_t peer_col0[8];  //   element (gi, gj) -> peer_sim[o][(gi - peer_row0[o]) * peer_pitch[o] + gj - peer_col0[o]]
  int64_t per_rows;
  int dist;
};

// further launches of one sweep_i8_launch call over other ranges of the tile list (same planes): own partial slots,
// an event recorded BEFORE each of them (= after the launch that precedes it)
struct SweepRange {
  int64_t t_begin, T;
  int G, max_seg;
  double* partial;
  cudaEvent_t before;
};

__host__ __device__ __forceinline__ void sweep_range(int64_t T, int64_t G, int64_t b, int64_t& t0, int64_t& t1) {
  t0 = T * b / G;
  t1 = T * (b + 1) / G;
}

// Symmetric sweep: column tile jt (64 wide, inside the 128-block jt / 2) only takes the row tiles 0 .. jt / 2,
// enumerated column-major like the full grid.  sym_offset(jt) = number of tiles before column tile jt.
__host__ __device__ __forceinline__ int64_t sym_offset(int64_t jt) {
  const int64_t q = jt >> 1, r = jt & 1;
  return (q + 1) * (q + r);
}
__host__ __device__ __forceinline__ int64_t sym_col_of(int64_t t) {  // largest jt with sym_offset(jt) <= t
  int64_t jt = 2 * (int64_t)sqrt((double)t);
  while (jt > 0 && sym_offset(jt) > t) --jt;
  while (sym_offset(jt + 1) <= t) ++jt;
  return jt;
}

// integer powers up to 64 are evaluated by repeated squaring (0: use pow())
__host__ __device__ __forceinline__ int small_int_power(double power) {
  return (power == (double)(int)power && power >= 1.0 && power <= 64.0) ? (int)power : 0;
}

#ifdef __CUDACC__
// s ** e by binary exponentiation for four values at once (same multiplication order for every value, so
// a symmetric pair of correlations gives bitwise equal adjacencies)
__device__ __forceinline__ void int_pow4(double (&s)[4], int e) {
  double r[4] = {1.0, 1.0, 1.0, 1.0};
#pragma unroll 1
  while (true) {
    if (e & 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) r[u] *= s[u];
    }
    e >>= 1;
    if (!e) break;
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] *= s[u];
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) s[u] = r[u];
}
#endif

// corr_i8.cu
int sweep_i8_launch(SweepParams prm, int64_t m, int64_t CT, int P, int G, int device, cudaStream_t st,
                    const SweepRange* more = nullptr, int n_more = 0);
bool corr_i8_enabled(int64_t m, int64_t n);

}  // namespace b200w
