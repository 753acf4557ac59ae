// K4, tensor-core flavour: L = A.A on tcgen05 (kind::i8, s8 x s8 -> s32 in TMEM) fed by TMA, with the
// TOM epilogue fused.  Replaces np `A @ A` + the elementwise tail of TomSimilarityFromAdj
// (PyWGCNA/wgcna.py:1145-1156).
//
// Arithmetic ("digit slices", an Ozaki-style split made for int8 tensor cores).  Every row i of the
// cleaned adjacency (NaN -> 0, diag -> 0, |a| <= 1) is scaled by a power of two 2^e_i so that its
// largest magnitude lies in [0.5, 1), rounded to a 38-bit fixed-point integer q_ij and written as S = 5
// balanced base-256 digits d^s_ij in [-128, 127]  (q = sum_s d^s 256^s).  Then
//     L_ij = sum_u a_iu a_ju = scale_i scale_j  sum_{s,t} 256^(s+t)  sum_u d^s_iu d^t_ju ,   scale = 2^-(e+38)
// where each inner sum is an s8 x s8 GEMM accumulated EXACTLY in int32 (A is symmetric to 1e-12 --
// checkAdjMat -- so the same digit planes serve as both operands, K-major).  Digit pairs are grouped
// by weight w = s + t; pairs with w < w_min (their total is below 2^-40 of the leading term per product)
// are dropped.  A "pass" accumulates pairs of one weight into one TMEM accumulator for at most 1023
// K-blocks of 128 (|sum| <= 1023 * 128 * 2^14 < 2^31: no overflow for ANY input), then the epilogue
// warps add acc * 256^(w - w_min) to a float64 running sum.  Integer accumulation has no rounding, so
// the result does not depend on the order of K-blocks, on the tile schedule or on the number of GPUs.
//
// Kernel structure (persistent, one CTA per SM, 320 threads; tom_i8_kernel<2> pairs the CTAs of two SMs
// into a cluster that works on one 256 x 256 tile with cta_group::2 MMAs, tom_i8_kernel<1> is the
// single-CTA 128 x 256 form):
//   warp 0     TMA producer: 6-stage ring of {row-operand tile 128 rows x 128 B, column-operand tile (half)
//              x 128 B}, SWIZZLE_128B; meets the producers of all CTAs at a global counter before every
//              digit-pair segment (split arrive / wait) so that co-resident CTAs keep sharing panels in L2
//   warp 1     MMA issuer (leader CTA of a pair): one lane issues tcgen05.mma K32, 4 per stage;
//              tcgen05.commit (multicast in a pair) frees the stage and, at the end of a pass, publishes
//              the accumulator
//   warp 2     also the TMEM allocator (512 columns = two s32 accumulators, double buffered)
//   warps 4-7  epilogue: tcgen05.ld 32 lanes x 32 columns, float64 running sum of the passes in an
//              L2-resident per-CTA scratch (two buffers)
//   warps 2, 3, 8, 9  finalizers: when a tile's last pass is in they apply (L + a) / (min(k_i,k_j) + 1 - a)
//              etc. and store the tile -- directly, mirrored (symmetric schedule: only tiles on and above the
//              diagonal are computed), into a peer's HBM (sharded run) or in condensed order -- and count
//              finished tiles per row block for a host that copies results out while the kernel runs
#include <cuda.h>

#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "tom_common.cuh"

namespace b200w {

constexpr int kS = B200W_I8_SLICES;  // digit planes stored
constexpr int kFixBits = 8 * kS - 2;  // 38
constexpr int kTM = 128, kTN = 256, kKB = 128;  // tile rows, tile cols, K bytes per stage
constexpr int kMaxSegs = 96, kMaxPasses = 48;
constexpr int kI8Threads = 320;  // warps: 0 TMA, 1 MMA, 2-3 + 8-9 finalizers (2 also owns TMEM), 4-7 epilogue
constexpr int kMaxKbPerPass = 1023;
constexpr int kCtasDefaultLead = 6;  // = ring stages of the 2-CTA kernel

struct Seg {
  int16_t s, t;    // digit planes of the left (row) and right (column) operand
  int32_t kb0, kb1;  // K-block range [kb0, kb1)
  int32_t pass;    // pass index this segment belongs to
};

struct I8Params {
  const double* A_rows;  // nrows x n original adjacency rows (for a_ij)
  const double* scale;   // op_rows
  const double* k;       // op_rows
  const double* scale_col;  // scale / k of the column operand (= scale / k unless A is asymmetric,
  const double* k_col;      //  TomColumnOperand)
  int32_t nan_mode;         // NaN rows / columns propagate (k = NaN marks them)
  // sharded run with the exchange overlapped: the operand rows of rank q (all planes, scale, k) are in place
  // once plane_flags[q] has reached plane_epoch (written on the copy stream after the copies from that peer)
  const unsigned int* plane_flags;
  unsigned int plane_epoch;
  int32_t self_rank;
  double* out;           // nrows x n
  double* scratch;       // gridDim.x * 128 * 256
  int64_t n, row0, nrows;
  int32_t tiles_m, tiles_n;
  const int2* tile_list;  // (tm, tn) per tile, in an order that keeps co-resident tiles on shared panels
  int32_t n_tiles;
  int32_t symmetric;      // every listed off-diagonal tile is also stored mirrored (TOM_ji := TOM_ij)
  int64_t tile_row_base;  // tile row index tm counts 256/128-row tiles from this row (row0, or 0 for global lists)
  int64_t per_rows;       // rows per rank: the mirrored entry (gj, gi) belongs to rank gj / per_rows ...
  double* peer_out[8];    // ... and is stored into that rank's row block (peer HBM over NVLink)
  int32_t nseg, npass;
  int32_t tom_type, tom_denom, out_mode, group;
  unsigned int* prog_count;           // optional: per row block, finalizer warps that finished a tile touching it
  const unsigned int* prog_expected;  //           value of prog_count at which the block's rows are complete
  unsigned int* prog_flag;            //           host-mapped flags, set to 1 when a block is complete
  int32_t prog_tiles;                 //           row tiles (of 256 rows) per row block
  int32_t fin_warps;       // finalizer warps per CTA: 4 (warps 2, 3, 8, 9; measured best) or 2 (warps 2, 3)
  int32_t sync_mode;       // 0 none, 1 grid barrier of the producers per tile, 2 per segment, 3 split (default)
  int32_t sync_lead;       // split barrier: K-blocks before a segment's end at which the next arrival is posted
  unsigned int* sync_ctr;  // zeroed before launch
  double final_scale;    // 256^w_min
  Seg seg[kMaxSegs];
  double pass_w[kMaxPasses];  // 256^(w - w_min)
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must end as a trapped launch, never as a hung GPU.  The bound is wall
// time (20 s on %globaltimer, checked every 4096 polls), not a poll count: under a profiler's
// instrumented replay a legitimate wait lasts orders of magnitude longer than in a plain run.
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
constexpr uint64_t kWaitLimitNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 4095u) == 0) {
      const uint64_t now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > kWaitLimitNs) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address
// [0,14) >> 4, LBO [16,30) (unused for swizzled K-major, 1), SBO [32,46) = 8 rows * 128 B = 1024 B,
// version [46,48) = 1, layout type [61,64) = 2 (SWIZZLE_128B).  Tiles are 1024-byte aligned.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format [4,6) = 2 (S32), a_format [7,10) = 1
// (signed int8), b_format [10,13) = 1, a/b major = K (0), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4.
__host__ __device__ constexpr uint32_t make_i8_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// Operand preparation: one CTA per row.  rmax -> e_i, 38-bit fixed point, 5 balanced digits per entry
// written to the 5 planes (rows zero padded to ldn); k_i = sum_j a_ij (wgcna.py:1146); scale_i.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tom_prepare_i8_kernel(const double* __restrict__ A_rows, int64_t n, int64_t ldn, int64_t op_rows,
                      int64_t row0, int8_t* __restrict__ planes, double* __restrict__ scale,
                      double* __restrict__ k, int nan_mode) {
  __shared__ double red[8];
  const int64_t i = row0 + blockIdx.x;
  const double* src = A_rows + (int64_t)blockIdx.x * n;
  double s = 0.0, mx = 0.0, nn = 0.0;
  for (int64_t j = threadIdx.x; j < n; j += 256) {
    if (nan_mode && i != j && isnan(src[j])) nn += 1.0;
    double a = tom_clean(src[j], i, j);
    s += a;
    mx = fmax(mx, fabs(a));
  }
  s = block_sum_256(s);
  if (nan_mode) {  // no nan_to_num upstream: numpy's plain row sum of a row with a NaN is NaN (wgcna.py:1146)
    nn = block_sum_256(nn);
    if (nn > 0.0) s = __longlong_as_double(0x7ff8000000000000LL);
  }
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
  int ex = 0;
  if (mx > 0.0) frexp(mx, &ex);  // mx = f * 2^ex, f in [0.5, 1)  ->  mx * 2^-ex in [0.5, 1)
  if (ex < -900) ex = -900;      // keeps 2^(38 - ex) finite; such rows contribute < 1e-270 anyway
  // up = 2^(kFixBits - ex): a * up is exact (power of two) and |a * up| <= 2^38
  const double up = ldexp(1.0, kFixBits - ex);
  if (threadIdx.x == 0) {
    k[i] = s;
    scale[i] = ldexp(1.0, ex - kFixBits);
  }
  const int64_t plane_stride = op_rows * ldn;
  int8_t* dst = planes + i * ldn;
  for (int64_t j4 = (int64_t)threadIdx.x * 4; j4 < ldn; j4 += 256 * 4) {
    uint32_t packed[kS];
#pragma unroll
    for (int p = 0; p < kS; ++p) packed[p] = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t j = j4 + e;
      long long q = 0;
      if (j < n) q = __double2ll_rn(tom_clean(src[j], i, j) * up);
#pragma unroll
      for (int p = 0; p < kS; ++p) {
        long long d = ((q + 128) & 255) - 128;  // balanced digit in [-128, 127]
        q = (q - d) >> 8;
        packed[p] |= ((uint32_t)(d & 255)) << (8 * e);
      }
    }
#pragma unroll
    for (int p = 0; p < kS; ++p)
      *reinterpret_cast<uint32_t*>(dst + p * plane_stride + j4) = packed[p];
  }
}

// ---------------------------------------------------------------------------------------------
// The contraction kernel.  kCtas = 1: one CTA per 128 x 256 tile.  kCtas = 2: a CTA pair (cluster of
// 2, tcgen05 cta_group::2) per 256 x 256 tile -- each CTA stages its own 128 rows of the row operand
// and its own 128-row half of the column operand, the leader CTA issues M256 N256 K32 MMAs that
// read both CTAs' shared memory and write both CTAs' TMEM.  Per SM that is 32 KB of operand per
// 512 tensor-pipe cycles instead of 48 KB.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // clears the CTA-pair bit of a shared::cluster address

template <int kCtas>
struct I8Cfg {
  static constexpr int kBRows = kTN / kCtas;                    // column-operand rows this CTA stages
  static constexpr int kStageBytes = (kTM + kBRows) * kKB;      // 48 KB / 32 KB
  static constexpr int kStages = kCtas == 1 ? 4 : 6;            // 192 KB of ring either way
  static constexpr int kTileRows = kTM * kCtas;                 // output rows per tile (CTA or pair)
};

template <int kStages>
struct __align__(8) I8Barriers {
  uint64_t full[kStages], empty[kStages], tmem_full[2], tmem_empty[2];
  uint64_t fin_full[2], fin_empty[2];  // epilogue warps -> finalizer warps hand-off of a finished tile's sums
  uint32_t tmem_base;
};

template <int kCtas>
__host__ __device__ constexpr size_t i8_smem_bytes() {
  return (size_t)I8Cfg<kCtas>::kStages * I8Cfg<kCtas>::kStageBytes + sizeof(I8Barriers<I8Cfg<kCtas>::kStages>) +
         1024;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// remote arrive on the barrier at the same offset in the even CTA of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                                int c1, int c2) {
  // executed by both CTAs; the transaction bytes are counted on the LEADER's barrier
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int kCtas>
__global__ void __launch_bounds__(kI8Threads, 1)
tom_i8_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmapB,
              const __grid_constant__ I8Params p) {
  using Cfg = I8Cfg<kCtas>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment (the dynamic smem base is the same in both CTAs of a
  // pair, so the aligned offsets match, as cta_group::2 requires)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  auto* bars = reinterpret_cast<I8Barriers<kStages>*>(smem + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = kCtas == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const int64_t unit = blockIdx.x / kCtas;           // CTA (pair) index
  const int64_t n_units = gridDim.x / kCtas;
  const int64_t n_tiles = p.n_tiles;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4 * kCtas);  // one arrive per epilogue warp of every CTA
      mbar_init(&bars->fin_full[a], 4);            // the 4 epilogue warps of this CTA
      mbar_init(&bars->fin_empty[a], p.fin_warps);  // the finalizer warps of this CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (kCtas == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                   "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                   "r"(512u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (kCtas == 1) __syncthreads(); else cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int64_t iters = (n_tiles + n_units - 1) / n_units;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      // Every CTA runs the same number of iterations so that the producers can keep in step: CTAs
      // that stream the same operand panels must do so within the L2 residency time, or each of them
      // fetches the panel from HBM again (measured: 525 GB of DRAM reads per launch without the
      // barrier against ~110 GB of unique panel bytes).
      unsigned int sync_seq = 0;
      for (int64_t it = 0; it < iters; ++it) {
        const int64_t tile = unit + it * n_units;
        const bool active = tile < n_tiles;
        int64_t tm = 0, tn = 0;
        if (active) {
          const int2 t2 = p.tile_list[tile];
          tm = t2.x;
          tn = t2.y;
        }
        const int i0 = (int)(p.tile_row_base + tm * Cfg::kTileRows + rank * kTM), j0 = (int)(tn * kTN);
        if (p.plane_flags != nullptr && active) {
          // sharded run with the exchange overlapped: the column operand of this tile belongs to the rank that owns
          // rows j0 ..; its rows (all digit planes, scale, k) are pulled by the copy engines while this kernel
          // runs (b200w_dev_tom_pull_planes), peer by peer in the order the tile list visits them, and
          // plane_flags[owner] reaches plane_epoch when they are in place.  Own rows need no wait.
          const int64_t owner = (int64_t)j0 / p.per_rows;
          if (owner != p.self_rank) {
            unsigned int v, spins = 0;
            uint64_t t0 = 0;
            while (true) {
              asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.plane_flags + owner) : "memory");
              if ((int)(v - p.plane_epoch) >= 0) break;
              if ((++spins & 1023u) == 0) {
                const uint64_t now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kWaitLimitNs) __trap();
              }
            }
            asm volatile("fence.proxy.async;" ::: "memory");  // the TMA reads below must see the copied data
          }
        }
        for (int sg = 0; sg < p.nseg; ++sg) {
          const Seg sgm = p.seg[sg];
          // sync_mode 3: split barrier -- the arrival for segment g was posted kLead K-blocks before the
          // end of segment g - 1 (below), so by the time a producer gets here most CTAs have arrived and the
          // ring does not drain; the skew between CTAs stays below one segment + kLead stages.
          const bool split = p.sync_mode == 3;
          if (split && sync_seq == 0) atomicAdd(p.sync_ctr, 1u);  // arrival for the very first segment
          if (p.sync_mode == 2 || split || (p.sync_mode == 1 && sg == 0)) {
            ++sync_seq;
            if (!split) {
              __threadfence();
              atomicAdd(p.sync_ctr, 1u);
            }
            const unsigned int target = sync_seq * gridDim.x;
            unsigned int v, spins = 0;
            uint64_t t0 = 0;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.sync_ctr) : "memory");
              if ((++spins & 4095u) == 0) {
                const uint64_t now = global_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kWaitLimitNs) __trap();
              }
            } while (v < target);
          }
          if (!active) {
            if (split) atomicAdd(p.sync_ctr, 1u);  // arrival for the next segment
            continue;
          }
          const int arrive_kb = (sgm.kb1 - p.sync_lead > sgm.kb0) ? sgm.kb1 - p.sync_lead : sgm.kb0;
          for (int kb = sgm.kb0; kb < sgm.kb1; ++kb) {
            if (split && kb == arrive_kb) atomicAdd(p.sync_ctr, 1u);  // arrival for the next segment
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* a_dst = smem + stage * kStageBytes;
            uint8_t* b_dst = a_dst + kTM * kKB;
            if constexpr (kCtas == 1) {
              mbar_expect_tx(&bars->full[stage], kStageBytes);
              tma_load_3d(a_dst, &tmap, &bars->full[stage], kb * kKB, i0, sgm.s);
              tma_load_3d(b_dst, &tmapB, &bars->full[stage], kb * kKB, j0, sgm.t);
              tma_load_3d(b_dst + 128 * kKB, &tmapB, &bars->full[stage], kb * kKB, j0 + 128, sgm.t);
            } else {
              if (leader) mbar_expect_tx(&bars->full[stage], 2 * kStageBytes);  // both CTAs' bytes
              tma_load_3d_2sm(a_dst, &tmap, &bars->full[stage], kb * kKB, i0, sgm.s);
              tma_load_3d_2sm(b_dst, &tmapB, &bars->full[stage], kb * kKB, j0 + (int)rank * 128, sgm.t);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ============================ MMA issuer (leader CTA only) ==============================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_i8_idesc(kTM * kCtas, kTN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t it = 0; it < iters; ++it) {
        if (unit + it * n_units >= n_tiles) break;
        int cur_pass = -1;
        bool first = true;
        for (int sg = 0; sg < p.nseg; ++sg) {
          const Seg sgm = p.seg[sg];
          if (sgm.pass != cur_pass) {  // a new pass: wait until the epilogue(s) drained this accumulator
            cur_pass = sgm.pass;
            mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            first = true;
          }
          const uint32_t tmem_d = tmem_base + (uint32_t)acc * kTN;
          for (int kb = sgm.kb0; kb < sgm.kb1; ++kb) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
            const uint64_t adesc = make_sw128_desc(a_addr);
            const uint64_t bdesc = make_sw128_desc(a_addr + kTM * kKB);
#pragma unroll
            for (int kk = 0; kk < kKB / 32; ++kk) {
              // K advance inside the 128-byte swizzle atom: +32 bytes on the start address (>> 4 = 2)
              if constexpr (kCtas == 1)
                tc_mma_i8(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, first ? 0u : 1u);
              else
                tc_mma_i8_2sm(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, first ? 0u : 1u);
              first = false;
            }
            // frees the smem stage (in both CTAs of a pair) when these MMAs retire
            if constexpr (kCtas == 1) tc_commit(&bars->empty[stage]); else tc_commit_2sm(&bars->empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          const bool pass_ends = (sg + 1 == p.nseg) || (p.seg[sg + 1].pass != cur_pass);
          if (pass_ends) {  // accumulator complete -> epilogue warps (of both CTAs)
            if constexpr (kCtas == 1) tc_commit(&bars->tmem_full[acc]); else tc_commit_2sm(&bars->tmem_full[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // ================ epilogue: TMEM accumulator -> float64 running sum in scratch ===============
    // Every pass is drained at once (coalesced scratch, lanes = rows) and the accumulator handed back
    // to the MMA warp.  The TOM formula and the n x n stores are NOT done here: measured at C2, doing
    // them before / after the release exposed 10 / 7 ms of 43, because these four warps were still busy
    // when the next accumulator filled.  The finished sums go to the two finalizer warps through a
    // double-buffered scratch tile.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int64_t done = 0;
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t tile = unit + it * n_units;
      if (tile >= n_tiles) break;
      const int buf = (int)(done & 1);
      double* scratch = p.scratch + ((int64_t)blockIdx.x * 2 + buf) * (kTM * kTN);
      mbar_wait(&bars->fin_empty[buf], (uint32_t)((done >> 1) & 1) ^ 1);  // finalizers are done with this buffer
      for (int ps = 0; ps < p.npass; ++ps) {
        mbar_wait(&bars->tmem_full[acc], acc_phase);
        tc_fence_after();
        const double w = p.pass_w[ps];
        const bool first = ps == 0;
#pragma unroll 1
        for (int c = 0; c < kTN / 32; ++c) {
          uint32_t r[32];
          tc_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTN + c * 32), r);
          double* sc = scratch + (int64_t)c * 32 * kTM + row;  // [c][col][row]: lanes = consecutive rows
#pragma unroll
          for (int col = 0; col < 32; ++col) {
            double v = (double)(int)r[col] * w;
            if (!first) v += sc[col * kTM];
            sc[col * kTM] = v;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kCtas == 1) mbar_arrive(&bars->tmem_empty[acc]); else mbar_arrive_leader(&bars->tmem_empty[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->fin_full[buf]);  // release: the sums of this tile are in scratch
      ++done;
    }
  } else if (warp - (warp < 4 ? 2 : 6) < p.fin_warps) {
    // ====== finalizers (warps 2, 3, 8, 9; 8 warps measured no better than 4, 2 warps 6 % slower): TOM formula + stores of a finished tile, off the MMA path ======
    // symmetric schedule: TOM_ji = TOM_ij, stored from the same registers; for a fixed column the 32
    // lanes of a warp hold 32 consecutive rows, so the mirrored store is one coalesced 256-byte line
    // (in a diagonal tile only the entries on and above the diagonal are stored, with their mirror);
    // the mirrored row gj may belong to another GPU: the store then goes to that GPU's HBM
    const int f = warp < 4 ? warp - 2 : warp - 6;  // warps 2, 3, 8, 9 -> finalizer 0..3
    const int nslab = p.fin_warps == 4 ? 1 : 2;  // 32-row slabs per finalizer
    const int slab0 = f * nslab;
    const bool mirror = p.symmetric != 0;
    int64_t done = 0;
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t tile = unit + it * n_units;
      if (tile >= n_tiles) break;
      const int2 t2 = p.tile_list[tile];
      const int64_t tm = t2.x, tn = t2.y;
      const int64_t j0 = tn * kTN;
      const bool diag_tile = p.tile_row_base + tm * Cfg::kTileRows == tn * kTN;
      const int buf = (int)(done & 1);
      const double* scratch = p.scratch + ((int64_t)blockIdx.x * 2 + buf) * (kTM * kTN);
      mbar_wait(&bars->fin_full[buf], (uint32_t)((done >> 1) & 1));
      // Latency, not bandwidth, is what this phase fights (one warp = one 32-row slab, little else to
      // switch to): all loads of a half chunk are issued up front and unconditionally (clamped addresses),
      // per-column scale / k come from one coalesced load per lane + shuffles, and only the stores are
      // predicated.  (With the loads inside the per-element branch this loop ran ~3700 cycles per element
      // and throttled the whole kernel through the scratch hand-off.)
#pragma unroll 1
      for (int slab = slab0; slab < slab0 + nslab; ++slab) {
        const int row = slab * 32 + lane;
        const int64_t gi = p.tile_row_base + tm * Cfg::kTileRows + rank * kTM + row;
        const bool row_ok = gi < p.row0 + p.nrows;
        const int64_t gic = row_ok ? gi : p.row0 + p.nrows - 1;  // clamped row for the loads
        const double ki = p.k[gic];
        const double si = p.scale[gic] * p.final_scale;
        const double* arow = p.A_rows + (gic - p.row0) * p.n;
        double* orow = p.out + (gic - p.row0) * p.n;
#pragma unroll 1
        for (int c = 0; c < kTN / 32; ++c) {
          const int64_t gjl = j0 + c * 32 + lane;
          const int64_t gjc = gjl < p.n ? gjl : p.n - 1;
          const double sj_l = p.scale_col[gjc], kj_l = p.k_col[gjc];  // column values: one per lane, shuffled below
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double* sc = scratch + ((int64_t)c * 32 + h * 16) * kTM + row;
            const int64_t gj0 = j0 + c * 32 + h * 16;
            double scv[16], av[16];
#pragma unroll
            for (int col = 0; col < 16; ++col) {
              scv[col] = sc[col * kTM];
              const int64_t gj = gj0 + col;
              av[col] = arow[gj < p.n ? gj : p.n - 1];
            }
#pragma unroll
            for (int col = 0; col < 16; ++col) {
              const int64_t gj = gj0 + col;
              const double sj = __shfl_sync(0xffffffffu, sj_l, h * 16 + col);
              const double kj = __shfl_sync(0xffffffffu, kj_l, h * 16 + col);
              const double L = scv[col] * (si * sj);
              const double a = tom_clean(av[col], gi, gj);
              double t = tom_value(L, a, ki, kj, gi == gj, p.tom_type, p.tom_denom, p.out_mode);
              if (p.nan_mode && gi != gj && (isnan(ki) || isnan(kj))) t = __longlong_as_double(0x7ff8000000000000LL);
              const bool ok = row_ok && gj < p.n;
              if (out_is_condensed(p.out_mode)) {  // entries above the diagonal only, squareform order
                if (ok && gj > gi) p.out[condensed_index(p.n, gi, gj)] = t;
                continue;
              }
              if (ok && !(mirror && diag_tile && gj < gi)) orow[gj] = t;
              if (ok && mirror && !(diag_tile && gj <= gi)) {
                const int64_t owner = gj / p.per_rows;
                p.peer_out[owner][(gj - owner * p.per_rows) * p.n + gi] = t;
              }
            }
          }
        }
      }
      if (p.prog_count != nullptr) {
        // progress for the host: when every finalizer warp of every tile that touches a row block has
        // passed here, the block's rows are final and the host starts copying them out while the rest of
        // the kernel is still running
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
          const int b0 = (int)(tm / p.prog_tiles), b1 = (int)(tn / p.prog_tiles);
          for (int b = b0;; b = b1) {
            const unsigned int old = atomicAdd(&p.prog_count[b], 1u);
            if (old + 1 == p.prog_expected[b]) {
              __threadfence_system();
              *(volatile unsigned int*)&p.prog_flag[b] = 1u;
            }
            if (b == b1) break;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->fin_empty[buf]);
      ++done;
    }
  }
  tc_fence_before();
  if constexpr (kCtas == 1) __syncthreads(); else cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (kCtas == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = (EncodeTiledFn)sym;
  return fn;
}

// true when Nsight Compute's injection libraries are loaded into this process (checked once)
static bool profiler_attached() {
  static const bool attached = [] {
    for (const char* v : {"CUDA_INJECTION64_PATH", "LD_PRELOAD"}) {
      const char* e = getenv(v);
      if (e && (strstr(e, "nsight") || strstr(e, "injection") || strstr(e, "Injection"))) return true;
    }
    FILE* f = fopen("/proc/self/maps", "r");
    if (!f) return false;
    char line[1024];
    bool hit = false;
    while (!hit && fgets(line, sizeof line, f))
      hit = strstr(line, "nsight-compute") || strstr(line, "libcuda-injection") || strstr(line, "TreeLauncher");
    fclose(f);
    return hit;
  }();
  return attached;
}

// Digit-pair schedule: pairs (s, t) with s + t >= w_min, grouped by weight (heaviest first), cut into
// passes of at most kMaxKbPerPass K-blocks.
static int build_schedule(I8Params& prm, int nkb, int w_min) {
  int nseg = 0, npass = 0;
  for (int w = 2 * (kS - 1); w >= w_min; --w) {
    int in_pass = 0;
    bool open = false;
    // pairs of one weight as mirror couples (s,t),(t,s): a pass that holds whole couples has a symmetric
    // accumulator, so L_ij and L_ji round identically; a new pass is opened before a couple that would
    // not fit (possible while 2 * nkb <= 1023, i.e. n <= 65408)
    for (int s = kS - 1; 2 * s >= w; --s) {
      const int t = w - s;
      if (t < 0 || t >= kS) continue;
      const int members = (s == t) ? 1 : 2;
      if (open && in_pass + members * nkb > kMaxKbPerPass && members * nkb <= kMaxKbPerPass) open = false;
      for (int mem = 0; mem < members; ++mem) {
        const int ps = mem == 0 ? s : t, pt = mem == 0 ? t : s;
        int kb = 0;
        while (kb < nkb) {
          if (!open || in_pass == kMaxKbPerPass) {
            if (npass == kMaxPasses) return -1;
            prm.pass_w[npass] = ldexp(1.0, 8 * (w - w_min));
            ++npass;
            in_pass = 0;
            open = true;
          }
          int take = nkb - kb;
          if (take > kMaxKbPerPass - in_pass) take = kMaxKbPerPass - in_pass;
          if (nseg == kMaxSegs) return -1;
          prm.seg[nseg].s = (int16_t)ps;
          prm.seg[nseg].t = (int16_t)pt;
          prm.seg[nseg].kb0 = kb;
          prm.seg[nseg].kb1 = kb + take;
          prm.seg[nseg].pass = npass - 1;
          ++nseg;
          kb += take;
          in_pass += take;
        }
      }
    }
  }
  prm.nseg = nseg;
  prm.npass = npass;
  prm.final_scale = ldexp(1.0, 8 * w_min);
  return 0;
}

}  // namespace b200w

using namespace b200w;

int b200w_tom_prepare_i8(const double* dA_rows, int64_t n, int64_t row0, int64_t nrows, void* d_operand,
                         int64_t op_rows, double* d_scale, double* d_k, cudaStream_t st, int nan_mode) {
  const int64_t ldn = b200w_padded_n(n);
  tom_prepare_i8_kernel<<<(unsigned)nrows, 256, 0, st>>>(dA_rows, n, ldn, op_rows, row0, (int8_t*)d_operand,
                                                         d_scale, d_k, nan_mode);
  B200W_LAUNCHED();
  return 0;
}

int b200w_tom_compute_i8(const double* dA_rows, const void* d_operand, int64_t op_rows, const double* d_scale,
                         const double* d_k, int64_t n, int64_t row0, int64_t nrows, int tom_type, int tom_denom,
                         int out_mode, double* d_out, cudaStream_t st, int world, int rank, int64_t per,
                         double* const* peer_out, b200w::TomProgress* progress, int w_min_arg,
                         const b200w::TomColumnOperand* col, const unsigned int* d_plane_flags,
                         unsigned int plane_epoch) {
  const int64_t ldn = b200w_padded_n(n);
  if (op_rows > 0x7fffffff || ldn > 0x7fffffff) return fail(B200W_E_UNSUPPORTED, "tom_i8: n too large");
  EncodeTiledFn encode = get_encode_tiled();
  if (!encode) return fail(B200W_E_CUDA, "tom_i8: cuTensorMapEncodeTiled not available from the driver");

  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)ldn, (cuuint64_t)op_rows, (cuuint64_t)kS};
  const cuuint64_t gstride[2] = {(cuuint64_t)ldn, (cuuint64_t)ldn * (cuuint64_t)op_rows};  // bytes, dims 1..2
  const cuuint32_t box[3] = {(cuuint32_t)kKB, 128u, 1u};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(d_operand), gdim, gstride, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(B200W_E_CUDA, "tom_i8: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  CUtensorMap tmapB = tmap;  // asymmetric adjacency: the column operand comes from the planes of A^T
  if (col != nullptr && col->operand != nullptr && col->operand != d_operand) {
    cr = encode(&tmapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(col->operand), gdim, gstride, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(B200W_E_CUDA, "tom_i8: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  }
  const bool asym_operand = col != nullptr && col->operand != nullptr && col->operand != d_operand;

  I8Params prm = {};  // per call: entry points may be driven from several host threads / for several devices
  prm.A_rows = dA_rows;
  prm.scale = d_scale;
  prm.k = d_k;
  prm.scale_col = (col && col->scale) ? col->scale : d_scale;
  prm.k_col = (col && col->k) ? col->k : d_k;
  prm.nan_mode = col ? col->nan_mode : 0;
  prm.plane_flags = d_plane_flags;
  prm.plane_epoch = plane_epoch;
  prm.self_rank = rank;
  prm.out = d_out;
  prm.n = n;
  prm.row0 = row0;
  prm.nrows = nrows;
  prm.tiles_m = (int32_t)((nrows + kTM - 1) / kTM);
  prm.tiles_n = (int32_t)((n + kTN - 1) / kTN);
  prm.tom_type = tom_type;
  prm.tom_denom = tom_denom;
  prm.out_mode = out_mode;
  prm.group = 8;
  if (const char* e = getenv("B200W_I8_GROUP")) prm.group = atoi(e) > 0 ? atoi(e) : 8;
  int w_min = w_min_arg > 0 ? w_min_arg : kS - 1;
  if (const char* e = getenv("B200W_I8_WMIN")) w_min = atoi(e);
  if (w_min < 0 || w_min > 2 * (kS - 1)) return fail(B200W_E_INVALID, "tom_i8: bad B200W_I8_WMIN");
  if (build_schedule(prm, (int)(ldn / kKB), w_min))
    return fail(B200W_E_UNSUPPORTED, "tom_i8: schedule too long for n = %lld", (long long)n);

  int ctas = 2;
  if (const char* e = getenv("B200W_I8_CTAS")) ctas = atoi(e) == 1 ? 1 : 2;
  prm.tiles_m = (int32_t)((nrows + kTM * ctas - 1) / (kTM * ctas));
  int dev = 0, sms = 0;
  B200W_CUDA(cudaGetDevice(&dev));
  B200W_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // Tile schedule.  Whole matrix on one GPU with square 256 x 256 tiles: TOM is symmetric (A is, to
  // 1e-12), so only the tiles on and above the diagonal are computed and mirrored -- half the MMAs.
  // The list is ordered in square super-blocks of G x G tiles so that the ~74 co-resident tiles touch
  // ~2G distinct operand panels.
  // Several GPUs (world > 1, peer_out given): the upper triangle of 256 x 256 tiles is dealt out over
  // the ranks -- tile {a, b}, a < b, goes to the owner of row tile a if a + b is even, else to the owner
  // of row tile b, oriented so that its direct rows are the rank's own; the mirrored half is stored
  // straight into the owning peer's row block (coalesced 256-byte lines over NVLink).
  const bool dist = world > 1 && peer_out != nullptr && ctas == 2;
  if (world > 8) return fail(B200W_E_UNSUPPORTED, "tom_i8: at most 8 ranks");
  if (dist && (per % 256 != 0 || row0 != rank * per))
    return fail(B200W_E_INVALID, "tom_i8: distributed mode needs per %% 256 == 0 and row0 == rank * per");
  bool symmetric = (ctas == 2) && ((row0 == 0 && nrows == n) || dist) && !asym_operand;
  if (const char* e = getenv("B200W_I8_SYMMETRIC")) symmetric = symmetric && (atoi(e) != 0 || dist);
  prm.tile_row_base = dist ? 0 : row0;
  prm.per_rows = dist ? per : op_rows;
  for (int r = 0; r < 8; ++r) prm.peer_out[r] = dist && r < world ? peer_out[r] : d_out;
  // the list depends only on (device, TM, TN, G, symmetric, world, rank, per): built once, then cached
  static std::map<std::vector<int>, std::pair<int2*, int>> tile_cache;
  static std::mutex tile_mu;
  std::lock_guard<std::mutex> tile_lock(tile_mu);  // (held to the end of the call: launches are asynchronous)
  const std::vector<int> key = {dev, prm.tiles_m, prm.tiles_n, prm.group, symmetric ? 1 : 0,
                                dist ? world : 1, dist ? rank : 0, dist ? (int)(per / 256) : 0};
  auto hit = tile_cache.find(key);
  if (hit == tile_cache.end()) {
    std::vector<int2> tiles;
    const int TM = prm.tiles_m, TN = prm.tiles_n;
    int G = prm.group > 0 ? prm.group : 8;
    if (dist) {
      const int T = TN, ptiles = (int)(per / 256);
      auto owner = [&](int t) { return t / ptiles; };
      for (int a = 0; a < T; ++a) {
        if (owner(a) == rank) tiles.push_back(make_int2(a, a));
        for (int b = a + 1; b < T; ++b) {
          const int oa = owner(a), ob = owner(b);
          const int o = (oa == ob) ? oa : (((a + b) & 1) == 0 ? oa : ob);
          if (o != rank) continue;
          tiles.push_back(oa == rank ? make_int2(a, b) : make_int2(b, a));
        }
      }
      // tiles whose column operand is local first, then the peers in the order their rows are pulled
      // (rank + 1, rank + 2, ...: b200w_dev_tom_pull_planes), inside a group the super-block order
      const int W = world;
      std::sort(tiles.begin(), tiles.end(), [G, W, rank, &owner](const int2& x, const int2& y) {
        const int dx = (owner(x.y) - rank + W) % W, dy = (owner(y.y) - rank + W) % W;
        const int kx[5] = {dx, x.y / G, x.x / G, x.y, x.x}, ky[5] = {dy, y.y / G, y.x / G, y.y, y.x};
        for (int i = 0; i < 5; ++i)
          if (kx[i] != ky[i]) return kx[i] < ky[i];
        return false;
      });
    } else if (symmetric) {
      // row-major over the super-blocks: when the block row bm is finished, rows [bm G 256, (bm+1) G 256)
      // are final (their direct tiles belong to this block row, their mirrored tiles to earlier ones)
      for (int bm = 0; bm * G < TM; ++bm)
        for (int bn = bm; bn * G < TN; ++bn)
          for (int tn = bn * G; tn < TN && tn < (bn + 1) * G; ++tn)
            for (int tm = bm * G; tm < TM && tm < (bm + 1) * G; ++tm)
              if (tm <= tn) tiles.push_back(make_int2(tm, tn));
    } else {
      for (int bm = 0; bm * G < TM; ++bm)
        for (int tn = 0; tn < TN; ++tn)
          for (int tm = bm * G; tm < TM && tm < (bm + 1) * G; ++tm) tiles.push_back(make_int2(tm, tn));
    }
    int2* d_tiles = nullptr;
    B200W_CUDA(cudaMalloc(&d_tiles, tiles.size() * sizeof(int2)));
    B200W_CUDA(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice));
    hit = tile_cache.emplace(key, std::make_pair(d_tiles, (int)tiles.size())).first;
  }
  prm.symmetric = symmetric ? 1 : 0;
  prm.n_tiles = hit->second.second;
  prm.tile_list = hit->second.first;
  const int64_t n_tiles = prm.n_tiles;
  const int64_t units = n_tiles < sms / ctas ? n_tiles : sms / ctas;
  const int grid = (int)(units * ctas);
  Scratch scratch;
  B200W_CUDA(scratch.alloc((size_t)grid * 2 * kTM * kTN * sizeof(double), st));  // two tiles per CTA
  prm.scratch = scratch.as<double>();
  prm.prog_count = nullptr;
  prm.prog_expected = nullptr;
  prm.prog_flag = nullptr;
  prm.prog_tiles = prm.group > 0 ? prm.group : 8;
  prm.fin_warps = 4;
  if (const char* e = getenv("B200W_I8_FIN")) prm.fin_warps = atoi(e) == 2 ? 2 : 4;
  prm.sync_mode = 3;
  if (const char* e = getenv("B200W_I8_SYNC")) prm.sync_mode = atoi(e);
  prm.sync_lead = kCtasDefaultLead;
  if (const char* e = getenv("B200W_I8_LEAD")) prm.sync_lead = atoi(e) > 0 ? atoi(e) : kCtasDefaultLead;
  if (progress != nullptr) {
    progress->nb = 0;
    if (symmetric && !dist && !out_is_condensed(out_mode)) {
      // expected finalizer-warp arrivals per row block = (warps per tile over the CTA pair) x (tiles that
      // touch the block through their direct rows tm or their mirrored rows tn)
      // one set per (host thread, device): a second device in the same process must not get device-0
      // pointers, and two host threads must not share counters
      struct ProgBufs {
        unsigned int* h_flags = nullptr;  // host-mapped, 1024 blocks
        unsigned int *d_flags = nullptr, *d_counts = nullptr, *d_expected = nullptr;
      };
      static thread_local std::map<int, ProgBufs> prog_bufs;
      ProgBufs& pb = prog_bufs[dev];
      if (!pb.h_flags) {
        B200W_CUDA(cudaHostAlloc((void**)&pb.h_flags, 1024 * sizeof(unsigned int), cudaHostAllocMapped));
        B200W_CUDA(cudaHostGetDevicePointer((void**)&pb.d_flags, pb.h_flags, 0));
        B200W_CUDA(cudaMalloc(&pb.d_counts, 1024 * sizeof(unsigned int)));
        B200W_CUDA(cudaMalloc(&pb.d_expected, 1024 * sizeof(unsigned int)));
      }
      unsigned int* const h_flags = pb.h_flags;
      unsigned int *const d_flags = pb.d_flags, *const d_counts = pb.d_counts, *const d_expected = pb.d_expected;
      const int Gp = prm.prog_tiles, nb = (prm.tiles_m + Gp - 1) / Gp;
      if (nb <= 1024) {
        std::vector<unsigned int> expect(nb, 0u);
        const unsigned int per_tile = (unsigned int)prm.fin_warps * 2u;
        for (int tm = 0; tm < prm.tiles_m; ++tm)
          for (int tn = tm; tn < prm.tiles_n; ++tn) {
            expect[tm / Gp] += per_tile;
            if (tn / Gp != tm / Gp) expect[tn / Gp] += per_tile;
          }
        for (int b = 0; b < nb; ++b) h_flags[b] = 0u;
        B200W_CUDA(cudaMemsetAsync(d_counts, 0, nb * sizeof(unsigned int), st));
        B200W_CUDA(cudaMemcpyAsync(d_expected, expect.data(), nb * sizeof(unsigned int), cudaMemcpyHostToDevice, st));
        B200W_CUDA(cudaStreamSynchronize(st));  // `expect` is pageable and local
        prm.prog_count = d_counts;
        prm.prog_expected = d_expected;
        prm.prog_flag = d_flags;
        progress->flags = h_flags;
        progress->nb = nb;
        progress->rows_per_block = (int64_t)Gp * 256;
      }
    }
  }
  Scratch ctr;
  B200W_CUDA(ctr.alloc(64, st));
  B200W_CUDA(cudaMemsetAsync(ctr.p, 0, 64, st));
  prm.sync_ctr = ctr.as<unsigned int>();

  const size_t smem = ctas == 1 ? i8_smem_bytes<1>() : i8_smem_bytes<2>();
  auto kern = ctas == 1 ? tom_i8_kernel<1> : tom_i8_kernel<2>;
  B200W_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool timed = getenv("B200W_TIME_KERNELS") != nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (timed) {
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
  }
  {
    // cooperative launch: every CTA is guaranteed co-resident, which the producer barrier needs
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kI8Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    // (Nsight Compute 2025.2 fails to replay a launch that is both cooperative and clustered, so the
    //  attribute is dropped when its injection libraries are mapped into the process -- or with
    //  B200W_I8_NOCOOP=1; the grid is one CTA per SM, so it is co-resident in practice)
    if (!getenv("B200W_I8_NOCOOP") && !profiler_attached()) {
      attr[na].id = cudaLaunchAttributeCooperative;
      attr[na].val.cooperative = 1;
      ++na;
    }
    if (ctas == 2) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmap, tmapB, prm);
    if (le != cudaSuccess) {
      return fail(B200W_E_CUDA, "tom_i8: launch failed (%s), grid %d x %d threads, %d-CTA groups, %zu B smem",
                  cudaGetErrorString(le), grid, kI8Threads, ctas, smem);
    }
  }
  B200W_LAUNCHED();
  if (timed) {
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    g_last_kernel_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  }
  return 0;
}
