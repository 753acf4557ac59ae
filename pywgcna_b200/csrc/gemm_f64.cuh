// FP64 tensor-core (DMMA, mma.sync.m16n8k8.f64) tile engine:  C[128 x BN] = P[128xK] * Q[BNxK]^T,
// BN = 128 (wide, 8 warps) or 64 (narrow, 4 warps), both operands K-major (row r of P / Q is K contiguous doubles, K % 16 == 0, 16-byte aligned rows).
// tcgen05 has no f64 kind, so the float64-exact products of this path (correlation for the sweep and
// the adjacency, and the TOM contraction at B200W_PREC_F64) run on the legacy-issue DMMA pipe.
//
// Wide: 256 threads = 8 warps arranged 2 (rows) x 4 (cols), narrow: 2 x 2; a warp owns a 64 x 32 sub-tile = 4 x 4 MMA
// tiles of 16 x 8, i.e. 64 accumulator doubles per thread.  Operands are staged through a
// cp.async ring of kStages slots of [128 rows][16 + 4 pad] doubles per operand (the +4 pad makes
// the 8-byte fragment loads of a half-warp hit 16 distinct bank pairs).
#pragma once
#include "common.cuh"

namespace b200w {

constexpr int kBM = 128, kBN = 128, kBK = 16, kLds = kBK + 4, kGemmThreads = 256;
constexpr int kStageDoubles = (kBM + kBN) * kLds;  // one ring slot, both operands

__host__ __device__ constexpr size_t gemm_smem_bytes(int stages) {
  return (size_t)stages * kStageDoubles * sizeof(double);
}

__device__ __forceinline__ void dmma_16x8x8(double (&d)[4], const double (&a)[4],
                                            const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// Element (mi, ni, c) of the per-thread accumulator sits at tile-local
//   row = wm*64 + mi*16 + (lane>>2) + ((c>>1) ? 8 : 0),  col = wn*32 + ni*8 + 2*(lane&3) + (c&1).
struct AccF64 {
  double v[4][4][4];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) v[i][j][c] = 0.0;
  }
};

// The engine comes in two widths: kWN = 4 (8 warps, 128 x 128 tile, one CTA per SM) and kWN = 2
// (4 warps, 128 x 64 tile, two CTAs per SM -- one CTA's MMAs run under the other one's prologue and
// epilogue; used where K is small and the epilogue long: the sweep and the adjacency).
// DMMA and ordinary FP64 arithmetic share one datapath on sm_100a (tools/f64_pipe: a DMMA.8x8x4 holds a
// sub-partition's FP64 pipe for 16 cycles, a DFMA warp instruction for 2.2, and mixed streams add up), so an
// FP64 epilogue is paid for in MMA time: epilogues here are kept to the arithmetic the reference needs.
template <int kWN>
struct GemmCfg {
  static constexpr int kBNt = 32 * kWN;           // tile columns
  static constexpr int kThreads = 64 * kWN;       // 2 x kWN warps
  static constexpr int kSlotDoubles = (kBM + kBNt) * kLds;
  __host__ __device__ static constexpr size_t smem_bytes(int stages) {
    return (size_t)stages * kSlotDoubles * sizeof(double);
  }
};

// Issue the cp.async copies of one ring slot.  Rows beyond the operand are clamped to its last row
// (their products are masked by the epilogues).
template <int kWN>
__device__ __forceinline__ void gemm_load_stage(double* slot, const double* __restrict__ P,
                                                const double* __restrict__ Q, int64_t ld,
                                                int64_t p_row0, int64_t p_rows, int64_t q_row0,
                                                int64_t q_rows, int64_t k0) {
  using Cfg = GemmCfg<kWN>;
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < kBM * 8 / Cfg::kThreads; ++i) {  // P: 128 rows x 8 chunks of 16 B
    int c = t + i * Cfg::kThreads;
    int r = c >> 3, kc = (c & 7) * 2;
    int64_t pr = p_row0 + r;
    if (pr >= p_rows) pr = p_rows - 1;
    cp_async16(slot + r * kLds + kc, P + pr * ld + k0 + kc);
  }
#pragma unroll
  for (int i = 0; i < Cfg::kBNt * 8 / Cfg::kThreads; ++i) {  // Q: kBNt rows x 8 chunks
    int c = t + i * Cfg::kThreads;
    int r = c >> 3, kc = (c & 7) * 2;
    int64_t qr = q_row0 + r;
    if (qr >= q_rows) qr = q_rows - 1;
    cp_async16(slot + kBM * kLds + r * kLds + kc, Q + qr * ld + k0 + kc);
  }
}

// acc += P[p_row0 : +128, 0:K] * Q[q_row0 : +128, 0:K]^T.  smem = kStages ring slots.  All threads
// of the CTA must call it; it ends with every thread past the last read of the ring.
template <int kStages, int kWN = 4>
__device__ __forceinline__ void gemm_f64_tile(AccF64& acc, double* smem,
                                              const double* __restrict__ P,
                                              const double* __restrict__ Q, int64_t ld, int64_t K,
                                              int64_t p_row0, int64_t p_rows, int64_t q_row0,
                                              int64_t q_rows) {
  constexpr int kSlot = GemmCfg<kWN>::kSlotDoubles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp / kWN, wn = warp % kWN, g = lane >> 2, tig = lane & 3;
  const int nk = (int)(K / kBK);

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nk)
      gemm_load_stage<kWN>(smem + s * kSlot, P, Q, ld, p_row0, p_rows, q_row0, q_rows, (int64_t)s * kBK);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    {
      int nxt = kt + kStages - 1;
      if (nxt < nk)
        gemm_load_stage<kWN>(smem + (nxt % kStages) * kSlot, P, Q, ld, p_row0, p_rows, q_row0, q_rows,
                             (int64_t)nxt * kBK);
      cp_async_commit();
    }
    const double* As = smem + (kt % kStages) * kSlot + (wm * 64 + g) * kLds + tig;
    const double* Bs = smem + (kt % kStages) * kSlot + kBM * kLds + (wn * 32 + g) * kLds + tig;
#pragma unroll
    for (int kk = 0; kk < kBK; kk += 8) {
      double a[4][4], b[4][2];
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const double* p = As + mi * 16 * kLds + kk;
        a[mi][0] = p[0];
        a[mi][1] = p[8 * kLds];
        a[mi][2] = p[4];
        a[mi][3] = p[8 * kLds + 4];
      }
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const double* p = Bs + ni * 8 * kLds + kk;
        b[ni][0] = p[0];
        b[ni][1] = p[4];
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma_16x8x8(acc.v[mi][ni], a[mi], b[ni]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

}  // namespace b200w
