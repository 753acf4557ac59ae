// K1 on the 5th-generation tensor cores: the standardise-and-correlate product Z Z^T of the soft-threshold sweep
// and of the adjacency (np.corrcoef at PyWGCNA/wgcna.py:882 and :1097) as tcgen05 kind::i8 MMAs fed by TMA, with
// the epilogues of corr.cu (clip -> transform -> power chain / s ** power, wgcna.py:884-916 and :1102-1111) fused.
//
// Arithmetic.  tcgen05 has no f64 kind, so -- as for the TOM contraction (tom_i8.cu) -- the product is formed
// exactly in integers: every standardised gene (a unit-norm row of Z, |z| <= 1) is scaled by 2^(46 - e_i), e_i
// the binary exponent of its largest magnitude, rounded to a 46-bit integer and cut into S = 6 balanced base-256
// digits (int8 planes).  cor_ij = 2^(e_i + e_j - 92) sum_{s,t} 256^(s+t) sum_k d^s_ik d^t_jk ; each inner sum is an
// s8 x s8 GEMM accumulated exactly in int32.  The 26 digit pairs of weight w = s + t >= 4 are kept.  (With w >= 5
// the dropped pair (2, 2) showed up: for strongly correlated genes and on the diagonal the digits of EQUAL planes
// are themselves correlated, so that product is a coherent sum ~ K * 5e3 * 256^4 -- measured 2e-13 at K = 200 and
// 2e-11 for a gene with one dominant sample; the other pairs of weight 4 add ~3e-14 of noise at small K.  Below
// weight 4 everything is under 1e-16.)  The pairs of one weight share one TMEM accumulator, so a tile holds SEVEN
// 128 x 64 accumulators (448 TMEM columns) and the epilogue combines them once:
// ((((((a10 * 256 + a9) * 256 + a8) ... ) * 256 + a5) * 256 + a4) * 2^32 * 2^(e_i + e_j - 92), every factor a power
// of two, so cor_ij and cor_ji are bitwise equal and do not depend on the tile schedule or on the number of GPUs.
// Measured against np.corrcoef: ~1e-15 (tests/test_gpu_round2.py).
//
// Kernel (persistent, one CTA per SM, 320 threads):
//   warp 8     TMA producer: per 64-byte K step ONE box {128 B, 128 rows, 3 plane pairs} of the row genes and one
//              {128 B, 64 rows, 3 plane pairs} of the column genes (the planes are stored pairwise interleaved, so a
//              box row is 128 bytes: with one 64-byte row per plane the kernel was bound by the TMA request rate),
//              SWIZZLE_128B, 2-stage ring of 72 KB -- all 26 pair products of a K step are issued from the same
//              staged bytes (36 B/clk of L2 traffic per SM instead of 125 with a stage per pair)
//   warp 9     TMEM allocation (512 columns) and MMA issue: 26 pairs x 2 tcgen05.mma M128 N64 K32 per stage
//   warps 0-7  epilogue, two warps per scheduler so that FP64 latencies overlap (with four warps the kernel ran
//              at 0.26 instructions per cycle and warp: profiles/r02_ncu_sweep_i8_v1.txt): every warp drains its
//              own 32 x 32 part of the seven accumulators (tcgen05.ld 32x32b.x16), combines, parks it in shared
//              memory and releases TMEM (the MMAs of the next tile run under the rest), then the epilogue of
//              corr.cu's sweep_kernel on the parked part: lane <-> column, power chain with the column sums in
//              registers (the number of powers is a template parameter: no branches inside the chain), second
//              pass with lane <-> row for tiles above the diagonal (symmetric schedule), optional similarity /
//              adjacency output
#include <cuda.h>

#include <vector>

#include "sweep_common.cuh"

namespace b200w {

constexpr int kZS = 6;                              // digit planes of Z
constexpr int kZBits = 8 * kZS - 2;                 // 46
constexpr int kZWmin = 4;                           // lowest digit-pair weight kept
constexpr int kZAcc = 2 * (kZS - 1) - kZWmin + 1;   // 7 accumulators, weights 4 .. 10
constexpr int kCKB = 64;                            // K bytes per stage
constexpr int kCStages = 2;
constexpr int kEpiWarps = 8, kTmaWarp = 8, kMmaWarp = 9;
constexpr int kCThreads = 320;                      // warps 0-7 epilogue, 8 TMA, 9 MMA
constexpr int kCAStage = kZS * kSweepBM * kCKB;     // 49152
constexpr int kCBStage = kZS * kSweepBN * kCKB;     // 24576
constexpr int kCStageBytes = kCAStage + kCBStage;   // 73728
constexpr int kWarpPitch = 33;                      // doubles; odd: row- and column-wise reads are conflict free
constexpr int kWarpPark = 32 * kWarpPitch;          // one epilogue warp's 32 x 32 sub-tile
constexpr int kParkDoubles = kEpiWarps * kWarpPark; // (>= 2 * kSweepMaxP * 128 for the row-sum exchange)

struct __align__(8) CorrBars {
  uint64_t full[kCStages], empty[kCStages], acc_full, acc_empty;
  uint32_t tmem_base;
};
constexpr size_t kCorrSmem = (size_t)kCStages * kCStageBytes + (size_t)kParkDoubles * 8 + sizeof(CorrBars) + 1024;

// ---- PTX wrappers (cta_group::1 only) ---------------------------------------------------------------------
namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded wait (20 s of wall time): a protocol bug ends as a trapped launch, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 4095u) == 0) {
      const uint64_t now = global_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 20ull * 1000ull * 1000ull * 1000ull) __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void ld_32x32_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address [0,14) >> 4,
// LBO [16,30) (unused for swizzled K-major, 1), SBO [32,46) = 8 rows * 128 B = 1024 B, version [46,48) = 1,
// layout type [61,64) = 2 (SWIZZLE_128B).  Tiles are 1024-byte aligned; a 32-byte K slice inside the 128-byte
// atom is addressed by adding its byte offset (0, 32, 64, 96) to the start address.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: c_format [4,6) = 2 (S32), a/b format [7,10) / [10,13) = 1 (signed int8), K-major both,
// n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
__host__ __device__ constexpr uint32_t make_i8_idesc(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
}  // namespace tc

// ---- operand preparation: one CTA per (padded) gene row of Z -------------------------------------------------
__global__ void __launch_bounds__(128)
z_prepare_i8_kernel(const double* __restrict__ Z, int64_t n, int64_t ldk, int64_t ldkb, int64_t rows_pad,
                    int8_t* __restrict__ planes, double* __restrict__ zscale) {
  __shared__ double red[4];
  // layout [plane pair q][row][K step of 64][128 B]: the 64 bytes of plane 2q, then those of plane 2q + 1 -- a TMA
  // box row is 128 bytes (half as many, twice as long requests as one 64-byte row per plane) and the two planes of a
  // pair sit in one SWIZZLE_128B atom, 64 bytes apart
  const int64_t i = blockIdx.x;
  const int64_t pair_stride = rows_pad * ldkb * 2;
  int8_t* dst = planes + i * ldkb * 2;
  double mx = 0.0;
  if (i < n)
    for (int64_t k = threadIdx.x; k < ldk; k += 128) mx = fmax(mx, fabs(Z[i * ldk + k]));
  for (int o = 16; o; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = fmax(fmax(red[0], red[1]), fmax(red[2], red[3]));
  int ex = 0;
  if (mx > 0.0) frexp(mx, &ex);  // mx = f 2^ex, f in [0.5, 1)
  if (ex < -900) ex = -900;
  const double up = ldexp(1.0, kZBits - ex);  // |z| up <= 2^46, exact scaling
  if (threadIdx.x == 0) zscale[i] = (i < n && mx > 0.0) ? ldexp(1.0, ex - kZBits) : 0.0;
  for (int64_t k4 = (int64_t)threadIdx.x * 4; k4 < ldkb; k4 += 128 * 4) {
    uint32_t packed[kZS];
#pragma unroll
    for (int p = 0; p < kZS; ++p) packed[p] = 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t k = k4 + e;
      long long q = 0;
      if (i < n && k < ldk) q = __double2ll_rn(Z[i * ldk + k] * up);
#pragma unroll
      for (int p = 0; p < kZS; ++p) {
        const long long d = ((q + 128) & 255) - 128;  // balanced digit in [-128, 127]
        q = (q - d) >> 8;
        packed[p] |= ((uint32_t)(d & 255)) << (8 * e);
      }
    }
#pragma unroll
    for (int p = 0; p < kZS; ++p)
      *reinterpret_cast<uint32_t*>(dst + (p >> 1) * pair_stride + (k4 >> 6) * 128 + (p & 1) * 64 + (k4 & 63)) = packed[p];
  }
}

// s ** e, the multiplication order of int_pow4 (bitwise the same adjacency as the DMMA route)
__device__ __forceinline__ double int_pow1(double s, int e) {
  double r = 1.0;
#pragma unroll 1
  while (true) {
    if (e & 1) r *= s;
    e >>= 1;
    if (!e) break;
    s *= s;
  }
  return r;
}

// the tile walk every role of the CTA repeats: column-major over the CT x RT grid (or its upper triangle)
struct TileWalk {
  int64_t t, t1, jt, it, RT;
  bool sym;
  const int2* list;
  __device__ __forceinline__ void init(const SweepParams& prm) {
    int64_t t0;
    sweep_range(prm.T, gridDim.x, blockIdx.x, t0, t1);
    list = prm.tile_list;
    if (list != nullptr) { t0 += prm.t_begin; t1 += prm.t_begin; }
    t = t0;
    sym = prm.symmetric != 0;
    RT = prm.RT;
    if (t0 >= t1) return;
    if (list != nullptr) {
      const int2 e = list[t0];
      it = e.x;
      jt = e.y;
    } else if (sym) {
      jt = sym_col_of(t0);
      it = t0 - sym_offset(jt);
    } else {
      jt = t0 / RT;
      it = t0 - jt * RT;
    }
  }
  __device__ __forceinline__ bool valid() const { return t < t1; }
  __device__ __forceinline__ void next() {
    ++t;
    if (list != nullptr) {
      if (t < t1) {
        const int2 e = list[t];
        it = e.x;
        jt = e.y;
      }
      return;
    }
    ++it;
    if (it >= (sym ? (jt >> 1) + 1 : RT)) { ++jt; it = 0; }
  }
};

// kP: powers evaluated (>= prm.P); kAllOne: every step of the chain is 1 (consecutive powers): no selects
template <int kP, bool kAllOne>
__global__ void __launch_bounds__(kCThreads, 1)
sweep_i8_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                const __grid_constant__ SweepParams prm, const int nkb) {
  constexpr int kBN = kSweepBN, kBM = kSweepBM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  double* park = reinterpret_cast<double*>(smem + kCStages * kCStageBytes);
  CorrBars* bars = reinterpret_cast<CorrBars*>(park + kParkDoubles);
  __shared__ uint8_t bad_col[kBN], bad_row[kBM];
  __shared__ double col_scale[kBN];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kMmaWarp && lane == 0) {
    for (int s = 0; s < kCStages; ++s) {
      tc::mbar_init(&bars->full[s], 1);
      tc::mbar_init(&bars->empty[s], 1);
    }
    tc::mbar_init(&bars->acc_full, 1);
    tc::mbar_init(&bars->acc_empty, kEpiWarps);  // lane 0 of every epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(&bars->tmem_base)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int64_t n = prm.n;

  if (warp == kTmaWarp) {
    // ================================= TMA producer =================================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmapB) : "memory");
      TileWalk w;
      w.init(prm);
      uint32_t step = 0;
      for (; w.valid(); w.next()) {
        const int i0 = (int)(w.it * kBM), j0 = (int)(prm.col0 + w.jt * kBN);
        for (int kb = 0; kb < nkb; ++kb, ++step) {
          const int stage = step & 1;
          tc::mbar_wait(&bars->empty[stage], ((step >> 1) & 1) ^ 1);
          uint8_t* a_dst = smem + stage * kCStageBytes;
          tc::mbar_expect_tx(&bars->full[stage], kCStageBytes);
          tc::tma_load_3d(a_dst, &tmapA, &bars->full[stage], kb * 128, i0, 0);
          tc::tma_load_3d(a_dst + kCAStage, &tmapB, &bars->full[stage], kb * 128, j0, 0);
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================================== MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_i8_idesc(kBM, kBN);
      TileWalk w;
      w.init(prm);
      uint32_t step = 0, tile_no = 0;
      long long dbg_empty = 0, dbg_full = 0, dbg_t = clock64();
      for (; w.valid(); w.next(), ++tile_no) {
        tc::mbar_wait(&bars->acc_empty, (tile_no & 1) ^ 1);  // the epilogue has drained the previous tile
        tc::fence_after();
        if (prm.debug) { const long long now = clock64(); dbg_empty += now - dbg_t; dbg_t = now; }
        for (int kb = 0; kb < nkb; ++kb, ++step) {
          const int stage = step & 1;
          if (prm.debug) dbg_t = clock64();
          tc::mbar_wait(&bars->full[stage], (step >> 1) & 1);
          tc::fence_after();
          if (prm.debug) { const long long now = clock64(); dbg_full += now - dbg_t; dbg_t = now; }
          const uint32_t a_addr = tc::smem_u32(smem + stage * kCStageBytes);
          const uint32_t b_addr = a_addr + kCAStage;
          uint32_t used = kb == 0 ? 0u : 0xffffffffu;  // accumulators that already hold a product of this tile
#pragma unroll
          for (int s = kZS - 1; s >= 0; --s) {
#pragma unroll
            for (int t = kZS - 1; t >= 0; --t) {
              if (s + t < kZWmin) continue;
              const int a = s + t - kZWmin;
              // plane p = pair p / 2 (a 128 rows x 128 B tile of its own), half p % 2 (64 bytes into the atom)
              const uint64_t adesc = tc::make_sw128_desc(a_addr + (s >> 1) * (kBM * 128)) + (uint64_t)((s & 1) * 4);
              const uint64_t bdesc = tc::make_sw128_desc(b_addr + (t >> 1) * (kBN * 128)) + (uint64_t)((t & 1) * 4);
              const uint32_t tmem_d = tmem_base + (uint32_t)(a * kBN);
#pragma unroll
              for (int kk = 0; kk < kCKB / 32; ++kk) {
                // K advance inside the swizzle atom: +32 bytes on the start address (>> 4 = 2)
                tc::mma_i8(tmem_d, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                           ((used >> a) & 1u) | (uint32_t)kk);
              }
              used |= 1u << a;
            }
          }
          tc::commit(&bars->empty[stage]);  // frees the stage when these MMAs retire
        }
        tc::commit(&bars->acc_full);  // all accumulators complete -> epilogue
        if (prm.debug) dbg_t = clock64();
      }
      if (prm.debug) { prm.debug[blockIdx.x * 8 + 5] = dbg_empty; prm.debug[blockIdx.x * 8 + 6] = dbg_full; }
    }
    __syncwarp();
  } else {
    // ========================= epilogue (warps 0-7, two per scheduler) =========================
    // warp = (wn, wm): rows 32 wm .. + 32 (its TMEM lane quarter), columns 32 wn .. + 32 -- a warp drains, parks
    // and post-processes ITS OWN 32 x 32 sub-tile, so drain -> chain needs no CTA-wide barrier
    const int wm = warp & 3, wn = warp >> 2;
    const int P = prm.P;
    const int64_t jend = prm.col0 + prm.ncols;
    double adj_power = prm.adj_power;
    int adj_ipower = prm.adj_ipower;
    if (prm.d_adj_power != nullptr) {
      adj_power = *prm.d_adj_power;
      adj_ipower = small_int_power(adj_power);
    }
    constexpr bool adj_mode = kP == 0;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    TileWalk w;
    w.init(prm);
    const int64_t jt_first = w.valid() ? w.jt : 0;
    int64_t cur_jt = -1;
    uint32_t tile_no = 0;
    double kcol[kP > 0 ? kP : 1];
#pragma unroll
    for (int p = 0; p < kP; ++p) kcol[p] = 0.0;
    bool one[kP > 0 ? kP : 1];
#pragma unroll
    for (int p = 0; p < kP; ++p) one[p] = prm.istep[p] == 1;  // powers past prm.P have step 1 and are not stored
    const int col = wn * 32 + lane;
    double* my = park + warp * kWarpPark;  // [row 0..31][col 0..31], pitch 33

    auto flush = [&](int64_t jt_done) {
      // column sums of a finished column tile -> this CTA's slot; the four warps that share a column add up
      // in a fixed order (deterministic, no atomics); rare: once per column tile and CTA
      double* dst = prm.partial + ((int64_t)blockIdx.x * prm.max_seg + (jt_done - jt_first)) * P * kBN;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        if (wm == q) {
#pragma unroll
          for (int p = 0; p < kP; ++p)
            if (p < P) dst[p * kBN + col] = (q == 0 ? 0.0 : dst[p * kBN + col]) + kcol[p];
        }
        tc::epi_barrier();
      }
#pragma unroll
      for (int p = 0; p < kP; ++p) kcol[p] = 0.0;
    };

    long long dbg[5] = {0, 0, 0, 0, 0}, dbg_t = clock64(), dbg7 = 0, dbg_t7 = 0;
    const bool dbg_on = prm.debug != nullptr && threadIdx.x == 0;
    auto mark = [&](int slot) {
      if (dbg_on) { const long long now = clock64(); dbg[slot] += now - dbg_t; dbg_t = now; }
    };
    for (; w.valid(); w.next(), ++tile_no) {
      const int64_t jt = w.jt, it = w.it, t = w.t;
      const int64_t j0 = prm.col0 + jt * kBN, i0 = it * kBM;
      // a tile off the diagonal block serves the genes of its rows too (symmetric schedules)
      const bool two_pass = prm.dist ? (it != ((prm.col0 + jt * kBN) >> 7)) : (w.sym && it < (jt >> 1));
      mark(4);
      tc::epi_barrier();  // every warp is done with bad_row, col_scale and the park of the previous tile
      if (jt != cur_jt) {
        if (kP > 0 && cur_jt >= 0) flush(cur_jt);
        if (threadIdx.x < kBN) {
          const int64_t j = j0 + threadIdx.x;
          bad_col[threadIdx.x] = (j < jend) ? prm.bad[j] : 1;
          // 2^32 = 256^4: the weight of the lowest accumulator
          col_scale[threadIdx.x] = (j < n) ? prm.zscale[j] * 4294967296.0 : 0.0;
        }
        cur_jt = jt;
      }
      if (threadIdx.x < kBM) {
        const int64_t i = i0 + threadIdx.x;
        bad_row[threadIdx.x] = (i < n) ? prm.bad[i] : 1;
      }
      tc::epi_barrier();
      // ---- drain: lane <-> row (TMEM lane), this warp's 32 columns of the accumulators combined in float64 ----
      {
        const int64_t gi = i0 + wm * 32 + lane;
        const double si = gi < n ? prm.zscale[gi] : 0.0;
        mark(0);  // barriers + per-tile setup
        tc::mbar_wait(&bars->acc_full, tile_no & 1);
        tc::fence_after();
        mark(1);  // waiting for the MMAs
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[kZAcc][16];
#pragma unroll
          for (int a = 0; a < kZAcc; ++a)
            tc::ld_32x32_x16(tmem_base + ((uint32_t)(wm * 32) << 16) + (uint32_t)(a * kBN + wn * 32 + c * 16), r[a]);
          tc::ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // Horner in 64-bit integers (exact), two conversions instead of seven: I2F.F64 is a slow instruction and
            // the drain sits on the critical path (the next tile's MMAs wait for it).  hi = weights 10 .. 7 (< 2^56),
            // lo = weights 6 .. 4 (< 2^48); the sum is rounded once per part like any float64 evaluation would be.
            long long hi = (long long)(int)r[kZAcc - 1][j];
#pragma unroll
            for (int a = kZAcc - 2; a >= 3; --a) hi = hi * 256 + (long long)(int)r[a][j];
            long long lo = (long long)(int)r[2][j];
            lo = lo * 256 + (long long)(int)r[1][j];
            lo = lo * 256 + (long long)(int)r[0][j];
            const double v = __ll2double_rn(hi) * 16777216.0 + __ll2double_rn(lo);
            my[lane * kWarpPitch + c * 16 + j] = v * (si * col_scale[wn * 32 + c * 16 + j]);  // powers of two: exact
          }
        }
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars->acc_empty);  // all 8 arrivals: the next tile's MMAs may start
        mark(2);  // drain
      }
      // ---- first pass: lane <-> column, the warp's 32 rows eight at a time ----
      // everything that does not depend on the element is hoisted: NaN flags of the 32 rows / columns as ballot
      // masks, validity as masks, the output pointer advanced by n per row
      const int64_t gj = j0 + col;
      const int64_t row_base = i0 + wm * 32;
      const bool cbad = bad_col[col];
      const int rows_valid = (int)((n - row_base) < 32 ? (n - row_base) : 32);  // may be <= 0
      const int diag_r = (int)((gj - row_base) >= 0 && (gj - row_base) < 32 ? gj - row_base : -1);
      const bool col_in = gj < jend;
      const unsigned rbad_mask = __ballot_sync(0xffffffffu, bad_row[wm * 32 + lane] != 0);
      const unsigned cbad_mask = __ballot_sync(0xffffffffu, cbad);
      const unsigned rvalid_mask = rows_valid >= 32 ? 0xffffffffu : (rows_valid <= 0 ? 0u : ((1u << rows_valid) - 1u));
      // (sharded: the diagonal block is covered by the transposed stores of its two column tiles)
      const bool write_direct = prm.sim != nullptr && (prm.dist ? two_pass : !prm.sim_transposed) && col_in;
      const bool keep_out = prm.sim != nullptr && prm.sim_transposed;
      double* out_ptr = prm.sim != nullptr ? prm.sim + row_base * n + gj : nullptr;
      int64_t out_pitch = n;
      if (prm.dist && prm.sim != nullptr) {
        // the tile's 128 rows belong to one rank (row blocks start on 256-row boundaries): element (gi, gj) goes into
        // that rank's block -- its own HBM, a peer's over NVLink, or a local rectangle staged for the copy engines
        const int64_t owner = i0 / prm.per_rows;
        out_pitch = prm.peer_pitch[owner];
        out_ptr = prm.peer_sim[owner] + (row_base - prm.peer_row0[owner]) * out_pitch + (gj - prm.peer_col0[owner]);
      }
      const int net_type = prm.net_type;

      // similarity of one element (wgcna.py:884-889 / :1102-1107) and what goes out with prm.sim: the similarity
      // (sweep) or s ** power (adjacency mode, :1111) the way WGCNA.adjacency forms it -- diagonal not forced, NaN
      // rows and columns kept.  Straight-line: the integer power is a fixed sequence of uniformly predicated
      // multiplications (same order as int_pow4: bitwise the adjacency of the DMMA route).
      auto out_value = [&](double sv) -> double {
        if (!adj_mode) return sv;
        if (adj_ipower <= 0) return pow(sv, adj_power);
        double r = 1.0, b = sv;
#pragma unroll
        for (int bit = 0; bit < 7; ++bit) {
          if ((adj_ipower >> bit) & 1) r *= b;
          if ((adj_ipower >> (bit + 1)) != 0) b *= b;
        }
        return r;
      };
      // clip to [-1, 1] (np.corrcoef) and the network transform WITHOUT float64 compares: on sm_100a DSETP /
      // DMNMX-style instructions hold the FP64 pipe for tens of cycles (the first version of this loader spent
      // 17 k of a tile's 33 k cycles in six of them per element) -- the sign and the "|c| >= 1" test are read off
      // the high word with integer instructions; c is finite by construction (integer sums times powers of two)
      double pre[16];
      auto similarity = [&](double c) -> double {
        int hi = __double2hiint(c);
        const int lo = __double2loint(c);
        const bool neg = hi < 0;
        const bool big = (hi & 0x7fffffff) >= 0x3ff00000;  // |c| >= 1  ->  +-1
        const int ahi = big ? 0x3ff00000 : (hi & 0x7fffffff);
        const int alo = big ? 0 : lo;
        if (net_type == B200W_NET_UNSIGNED) return __hiloint2double(ahi, alo);                  // |c|
        if (net_type == B200W_NET_SIGNED_HYBRID) return neg ? 0.0 : __hiloint2double(ahi, alo);  // max(c, 0)
        return (1.0 + __hiloint2double(neg ? (ahi | 0x80000000) : ahi, alo)) / 2.0;              // (1 + c) / 2
      };

      // Shared-memory loads issued one by one stall for hundreds of cycles while the tensor core streams its operands
      // out of shared memory (N = 64 MMAs read 97 B/clk; measured: this pass took 17 k cycles inside the MMA window
      // and 6 k outside), so 16 rows are fetched at once and their latencies overlap.
#pragma unroll 1
      for (int h0 = 0; h0 < 32; h0 += 16) {
#pragma unroll
        for (int q = 0; q < 16; ++q) pre[q] = my[(h0 + q) * kWarpPitch + lane];
#pragma unroll
      for (int g8 = 0; g8 < 16; g8 += 8) {
        const int r0 = h0 + g8;
        double s1[8], s2[8], cur[8];
        if (dbg_on) dbg_t7 = clock64();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = r0 + u;
          double sv = similarity(pre[g8 + u]);
          const bool nan_pair = cbad || ((rbad_mask >> r) & 1u);
          const bool valid = (rvalid_mask >> r) & 1u;
          if (prm.sim != nullptr) {
            const double s_out = nan_pair ? qnan : out_value(sv);
            if (keep_out) my[r * kWarpPitch + lane] = s_out;  // stored by the transposed pass below
            if (write_direct && valid) out_ptr[(int64_t)r * out_pitch] = s_out;
          }
          // chain start: the diagonal counts as 1 even for a NaN gene (wgcna.py:897), NaN pairs are skipped by
          // nansum (:915), rows past the matrix contribute nothing
          double c0 = 1.0;
          if (r == diag_r) sv = 1.0;
          else if (nan_pair) { sv = 0.0; c0 = 0.0; }
          if (!valid) { sv = 0.0; c0 = 0.0; }
          s1[u] = sv;
          s2[u] = sv * sv;
          cur[u] = c0;
        }
        if (dbg_on) {  // loader share of the first pass (slot 7)
          const long long now = clock64();
          dbg7 += now - dbg_t7;
        }
#pragma unroll
        for (int p = 0; p < kP; ++p) {  // wgcna.py:914  corxCur = corxPrev * corx ** step
#pragma unroll
          for (int u = 0; u < 8; ++u) cur[u] *= (kAllOne || one[p]) ? s1[u] : s2[u];
          kcol[p] += ((cur[0] + cur[1]) + (cur[2] + cur[3])) + ((cur[4] + cur[5]) + (cur[6] + cur[7]));
        }
      }
      }
      mark(3);  // first pass
      if (prm.sim != nullptr && prm.sim_transposed) {  // lane <-> row of the sub-tile: 256-byte row segments
        // (before the second pass: its row-sum exchange reuses the parked tile)
        __syncwarp();
        const int64_t gi = i0 + wm * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < 32; ++c) {
          const int64_t gjc = j0 + wn * 32 + c;
          if (gjc >= jend) break;  // warp-uniform
          if (gi < n) prm.sim[(gjc - prm.col0) * n + gi] = my[lane * kWarpPitch + c];
        }
      }
      if (two_pass) {
        // lane <-> row: row sums over the warp's 32 columns, eight at a time, for the genes of the tile's rows
        // (their column tiles never visit these rows), and the mirrored half of the output
        double kr[kP > 0 ? kP : 1];
#pragma unroll
        for (int p = 0; p < kP; ++p) kr[p] = 0.0;
        const int64_t gi = row_base + lane;
        const bool rbad = (rbad_mask >> lane) & 1u;
        const bool row_in = gi < n;
        const int64_t col_base = j0 + wn * 32;
        const int cols_valid = (int)((jend - col_base) < 32 ? (jend - col_base) : 32);
        const unsigned cvalid_mask = cols_valid >= 32 ? 0xffffffffu : (cols_valid <= 0 ? 0u : ((1u << cols_valid) - 1u));
        // (sharded: both orientations are already written by the first pass and the transposed pass)
        double* mir_ptr = (prm.sim != nullptr && !prm.dist) ? prm.sim + col_base * n + gi : nullptr;
        __syncwarp();
#pragma unroll 1
        for (int h0 = 0; h0 < 32; h0 += 16) {
#pragma unroll
          for (int q = 0; q < 16; ++q) pre[q] = my[lane * kWarpPitch + h0 + q];
#pragma unroll
        for (int g8 = 0; g8 < 16; g8 += 8) {
          const int c0 = h0 + g8;
          double s1[8], s2[8], cur[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int cx = c0 + u;
            // (with keep_out the first pass has left the similarity itself in the tile, NaN for NaN pairs)
            double sv = keep_out ? pre[g8 + u] : similarity(pre[g8 + u]);
            const bool nan_pair = rbad || ((cbad_mask >> cx) & 1u);
            const bool inside = row_in && ((cvalid_mask >> cx) & 1u);
            if (mir_ptr != nullptr && inside) mir_ptr[(int64_t)cx * n] = nan_pair ? qnan : out_value(sv);
            double c0v = 1.0;
            if (nan_pair || !inside) { sv = 0.0; c0v = 0.0; }
            s1[u] = sv;
            s2[u] = sv * sv;
            cur[u] = c0v;
          }
#pragma unroll
          for (int p = 0; p < kP; ++p) {
#pragma unroll
            for (int u = 0; u < 8; ++u) cur[u] *= (kAllOne || one[p]) ? s1[u] : s2[u];
            kr[p] += ((cur[0] + cur[1]) + (cur[2] + cur[3])) + ((cur[4] + cur[5]) + (cur[6] + cur[7]));
          }
        }
        }
        if (kP > 0) {
          // the two warps that share rows add up through the (now free) parking space: [wn][p][row]
          tc::epi_barrier();
#pragma unroll
          for (int p = 0; p < kP; ++p)
            if (p < P) park[(wn * P + p) * kBM + wm * 32 + lane] = kr[p];
          tc::epi_barrier();
          if (threadIdx.x < kBM) {
            const int row = threadIdx.x;
            double* dst = prm.rowpart + t * P * kBM;
            for (int p = 0; p < P; ++p) dst[p * kBM + row] = park[p * kBM + row] + park[(P + p) * kBM + row];
          }
        }
      }
    }
    mark(4);  // second pass, row-sum exchange, transposed stores
    if (kP > 0 && cur_jt >= 0) {
      tc::epi_barrier();
      flush(cur_jt);
    }
    if (dbg_on)
      for (int i = 0; i < 5; ++i) prm.debug[blockIdx.x * 8 + i] = dbg[i];
    if (dbg_on) prm.debug[blockIdx.x * 8 + 7] = dbg7;
  }
  tc::fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn corr_encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      sym = nullptr;
    return (EncodeTiledFn)sym;
  }();
  return fn;
}

static std::atomic<int> g_corr_precision{-1};  // -1: environment / default

bool corr_i8_enabled(int64_t m, int64_t n) {
  int prec = g_corr_precision.load();
  if (prec < 0) {
    const char* e = getenv("B200WGCNA_CORR_PRECISION");
    prec = (e && (e[0] == 'f' || e[0] == 'F')) ? B200W_PREC_F64 : B200W_PREC_I8;
  }
  // per accumulator |sum| <= K * 2^14 * (pairs of one weight <= 6) must stay below 2^31
  const int64_t ldkb = round_up(m, kCKB);
  return prec != B200W_PREC_F64 && ldkb * 6 * 16384 < 2147483647LL && n >= 1;  // (weight 5 has the most pairs: 6)
}

// One launch of the tcgen05 sweep / adjacency engine.  prm is filled by the caller as for corr.cu's sweep_kernel
// (partial / rowpart / T / max_seg / symmetric ...); planes and scales are made here, stream ordered.
int sweep_i8_launch(SweepParams prm, int64_t m, int64_t CT, int P, int G, int device, cudaStream_t st,
                    const SweepRange* more, int n_more) {
  EncodeTiledFn encode = corr_encode_tiled();
  if (!encode) return fail(B200W_E_CUDA, "corr_i8: cuTensorMapEncodeTiled not available from the driver");
  const int64_t n = prm.n;
  const int64_t ldkb = round_up(m, kCKB), rows_pad = round_up(n, 128);
  if (rows_pad > 0x7fffffff || ldkb > 0x7fffffff) return fail(B200W_E_UNSUPPORTED, "corr_i8: problem too large");
  Scratch planes, zscale;
  B200W_CUDA(planes.alloc((size_t)kZS * rows_pad * ldkb, st));
  B200W_CUDA(zscale.alloc((size_t)rows_pad * sizeof(double), st));
  z_prepare_i8_kernel<<<(unsigned)rows_pad, 128, 0, st>>>(prm.Z, n, prm.ldk, ldkb, rows_pad, planes.as<int8_t>(),
                                                           zscale.as<double>());
  B200W_LAUNCHED();
  prm.zscale = zscale.as<double>();
  prm.debug = nullptr;
  Scratch dbg;
  static const bool debug_on = getenv("B200W_CORR_DEBUG") != nullptr;
  if (debug_on) {
    B200W_CUDA(dbg.alloc((size_t)G * 8 * sizeof(long long), st));
    B200W_CUDA(cudaMemsetAsync(dbg.p, 0, (size_t)G * 8 * sizeof(long long), st));
    prm.debug = dbg.as<long long>();
  }

  CUtensorMap tmapA, tmapB;
  // [pair][row][2 ldkb bytes]: a box row = one K step of one gene = 64 B of plane 2q + 64 B of plane 2q + 1
  const cuuint64_t gdim[3] = {(cuuint64_t)ldkb * 2, (cuuint64_t)rows_pad, (cuuint64_t)(kZS / 2)};
  const cuuint64_t gstride[2] = {(cuuint64_t)ldkb * 2, (cuuint64_t)ldkb * 2 * (cuuint64_t)rows_pad};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const cuuint32_t boxA[3] = {128u, (cuuint32_t)kSweepBM, (cuuint32_t)(kZS / 2)};
  const cuuint32_t boxB[3] = {128u, (cuuint32_t)kSweepBN, (cuuint32_t)(kZS / 2)};
  CUresult cr = encode(&tmapA, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, planes.p, gdim, gstride, boxA, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr == CUDA_SUCCESS)
    cr = encode(&tmapB, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, planes.p, gdim, gstride, boxB, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(B200W_E_CUDA, "corr_i8: cuTensorMapEncodeTiled failed (%d)", (int)cr);
  // the power count is a template parameter (no branches inside the chain); powers past prm.P multiply by s and
  // are not stored
  auto launch = [&](auto kern) -> int {
    B200W_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCorrSmem));
    if (prm.T > 0) {
      kern<<<G, kCThreads, kCorrSmem, st>>>(tmapA, tmapB, prm, (int)(ldkb / kCKB));
      B200W_LAUNCHED();
    }
    for (int i = 0; i < n_more; ++i) {  // further ranges of the tile list on the same planes
      const SweepRange& r = more[i];
      if (r.before) B200W_CUDA(cudaEventRecord(r.before, st));
      SweepParams p2 = prm;
      p2.t_begin = r.t_begin;
      p2.T = r.T;
      p2.max_seg = r.max_seg;
      p2.partial = r.partial;
      if (p2.T > 0) {
        kern<<<r.G, kCThreads, kCorrSmem, st>>>(tmapA, tmapB, p2, (int)(ldkb / kCKB));
        B200W_LAUNCHED();
      }
    }
    return 0;
  };
  bool all_one = true;
  for (int p = 0; p < prm.P; ++p) all_one = all_one && prm.istep[p] == 1;
  int rc;
  if (prm.P == 0) rc = launch(sweep_i8_kernel<0, true>);
  else if (prm.P <= 2) rc = launch(sweep_i8_kernel<2, false>);
  else if (prm.P <= 10) rc = launch(sweep_i8_kernel<10, false>);
  else if (prm.P <= 15) rc = launch(sweep_i8_kernel<15, false>);
  else if (all_one) rc = launch(sweep_i8_kernel<kSweepMaxP, true>);
  else rc = launch(sweep_i8_kernel<kSweepMaxP, false>);
  if (rc) return rc;
  if (debug_on) {  // development aid: where the cycles of epilogue warp 0 and of the MMA warp go
    std::vector<long long> h((size_t)G * 8);
    B200W_CUDA(cudaMemcpyAsync(h.data(), dbg.p, h.size() * sizeof(long long), cudaMemcpyDeviceToHost, st));
    B200W_CUDA(cudaStreamSynchronize(st));
    double a[8] = {0};
    for (int b = 0; b < G; ++b)
      for (int i = 0; i < 8; ++i) a[i] += (double)h[(size_t)b * 8 + i] / G;
    const double tiles = (double)prm.T / G;
    fprintf(stderr, "[b200w] sweep_i8 P=%d: per tile (cycles): setup+barriers %.0f  wait MMA %.0f  drain %.0f  pass1 %.0f  "
                    "(loader %.0f)  pass2+rest %.0f | MMA warp: wait drained %.0f  wait TMA %.0f  (%.0f tiles per CTA)\n",
            prm.P, a[0] / tiles, a[1] / tiles, a[2] / tiles, a[3] / tiles, a[7] / tiles, a[4] / tiles, a[5] / tiles,
            a[6] / tiles, tiles);
  }
  (void)CT;
  (void)P;
  (void)device;
  return 0;
}

}  // namespace b200w

extern "C" int b200w_set_correlation_precision(int precision) {
  if (precision != B200W_PREC_AUTO && precision != B200W_PREC_F64 && precision != B200W_PREC_I8)
    return b200w::fail(B200W_E_INVALID, "set_correlation_precision: B200W_PREC_AUTO, _F64 or _I8");
  b200w::g_corr_precision.store(precision == B200W_PREC_AUTO ? -1 : precision);
  return 0;
}
