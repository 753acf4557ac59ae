// Issue-rate microbenchmark of tcgen05.mma on B200: one CTA per SM, one thread issues `iters` MMAs
// back to back on fixed (garbage) shared-memory operands, commit, wait; cycles per MMA from clock64.
// Measures the tensor-pipe floor per instruction for kind::i8 / kind::f8f6f4 / kind::f16 at M = 128
// (cta_group::1) and N in {64, 128, 256} -- the denominator for the roofline of the TOM kernel.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KIND>  // 0 = i8, 1 = f8f6f4 (e4m3), 2 = f16 (bf16)
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == 1)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(256, 1) rate_kernel(int iters, int N, long long* out, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t ring[8];
  __shared__ uint64_t fullb[8], emptyb[8];
  __shared__ uint64_t never;
  __shared__ volatile int done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 112 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x01010101u;
  if (warp == 0 && lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&ring[i])), "r"(1));
    for (int i = 0; i < 8; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&fullb[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&emptyb[i])), "r"(1));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&never)), "r"(1));
    done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (warp == 0 && lane == 0) {
    // instruction descriptor: c_format [4,6), a_format [7,10), b_format [10,13), n [17,23), m [24,29)
    uint32_t idesc;
    if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10);        // S32 <- s8 x s8
    else if (KIND == 1) idesc = (1u << 4) | (0u << 7) | (0u << 10);   // F32 <- e4m3 x e4m3
    else idesc = (1u << 4) | (1u << 7) | (1u << 10);                  // F32 <- bf16 x bf16
    idesc |= ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc = make_sw128_desc(smem_u32(smem));
    const uint64_t bdesc = make_sw128_desc(smem_u32(smem) + 16 * 1024);
    long long t0 = clock64();
    if (mode & 8) {
      const int S = 6;
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < iters; i += 4) {
        uint32_t ok = 0;
        while (!ok)
          asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&fullb[stage])), "r"(phase) : "memory");
        if (!(mode & 16)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t off = (uint64_t)((stage % 3) * (32 * 1024 >> 4));
        for (int k = 0; k < 4; ++k)
          mma<KIND>(tmem, adesc + off + (uint64_t)(k * 2), bdesc + off + (uint64_t)(k * 2), idesc, (i | k) ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&emptyb[stage])) : "memory");
        if (++stage == S) { stage = 0; phase ^= 1; }
      }
    } else
    for (int i = 0; i < iters; ++i) {
      const uint64_t off = (mode & 4) ? (uint64_t)(((i >> 2) % 3) * (32 * 1024 >> 4)) : 0ull;
      mma<KIND>(tmem, adesc + off + (uint64_t)((i & 3) * 2), bdesc + off + (uint64_t)((i & 3) * 2), idesc, i ? 1u : 0u);
      if ((mode & 1) && (i & 3) == 3)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ring[(i >> 2) & 7])) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
    done = 1;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&never)) : "memory");
  } else if (warp == 2 && lane == 0 && (mode & 8)) {
    // producer: waits for a free slot, "fills" it (no data movement), publishes it
    const int S = 6;
    int stage = 0; uint32_t phase = 0;
    for (int i = 0; i < iters; i += 4) {
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&emptyb[stage])), "r"(phase ^ 1) : "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&fullb[stage])) : "memory");
      if (++stage == S) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4 && (mode & 2)) {
    // polling warps, as the epilogue warps of the TOM kernel wait for an accumulator
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&never)), "r"(0) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

template <int KIND>
void run(const char* name, int kbytes_per_mma_k, int mode) {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  for (int grid : {148})
    for (int N : {64, 128, 256}) {
      const int iters = 20000;
      rate_kernel<KIND><<<grid, 256, 128 * 1024>>>(1000, N, d, mode);  // warm
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      rate_kernel<KIND><<<grid, 256, 128 * 1024>>>(iters, N, d, mode);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      long long cyc = 0;
      cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double macs = (double)grid * iters * 128.0 * N * kbytes_per_mma_k;
      printf("%-8s grid %3d M128 N%3d: %7.1f cycles/MMA (ideal %3d at 8192 8-bit MAC/clk/SM), %8.1f T MAC-op/s x2 = %7.1f TOP/s  [%s]\n",
             name, grid, N, (double)cyc / iters, KIND == 2 ? N / 2 * 16 / 16 : N / 2, macs / (ms * 1e-3) / 1e12,
             2 * macs / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
    }
  cudaFree(d);
}

int main() {
  for (int mode : {0, 8, 24, 10}) {
    printf("--- mode %d (8 = full/empty ring with a producer thread + fence per stage, 24 = same without the fence, 10 = ring + polling warps)\n", mode);
    run<0>("i8", 32, mode);
  }
  run<2>("bf16", 16, 0);
  return 0;
}
