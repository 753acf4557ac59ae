// Does DMMA (mma.sync f64) share its datapath with ordinary FP64 arithmetic (DFMA/DMUL/DADD) on sm_100a?
// Each SM sub-partition gets `wd` warps issuing independent DMMA.8x8x4 and `wf` warps issuing independent
// DFMA chains; the time of the mixed run against the two pure runs answers the question, and the pure
// runs give the per-warp issue limits (1 warp vs 2 warps per sub-partition).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f64_pipe f64_pipe.cu ; run: ./f64_pipe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// warps [0, 4*wd) issue DMMA, warps [4*wd, 4*(wd+wf)) issue DFMA; warp w sits on sub-partition w % 4
__global__ void mix_kernel(int wd, int iters_d, int iters_f, double* out, double seed) {
  const int warp = threadIdx.x >> 5;
  if (warp < 4 * wd) {
    double acc[16][2];
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = seed + threadIdx.x, b = seed * 0.5;
    for (int it = 0; it < iters_d; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) dmma884(acc[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 12345.678) out[threadIdx.x] = s;
  } else {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = seed + i;
    const double m = 1.0 + seed * 1e-9, q = seed * 1e-12;
    for (int it = 0; it < iters_f; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) c[i] = fma(c[i], m, q);
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[threadIdx.x] = s;
  }
}

static float run(int wd, int wf, int itd, int itf, double* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int threads = 128 * (wd + wf);
  mix_kernel<<<148, threads>>>(wd, itd, itf, out, 1.0);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  mix_kernel<<<148, threads>>>(wd, itd, itf, out, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* out; cudaMalloc(&out, 1 << 16);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int itd = 20000, itf = 80000;  // 16 DMMA resp. 16 DFMA per iteration
  struct { int wd, wf; } cfg[] = {{1,0},{2,0},{4,0},{0,1},{0,2},{0,4},{1,1},{2,2},{1,2},{2,1}};
  for (auto c : cfg) {
    float ms = run(c.wd, c.wf, itd, itf, out);
    double dmma_per_clk = c.wd ? (double)c.wd * itd * 16 / (ms * 1e-3 * clk * 1e3) : 0;  // per sub-partition
    double dfma_per_clk = c.wf ? (double)c.wf * itf * 16 / (ms * 1e-3 * clk * 1e3) : 0;
    printf("dmma warps/smsp %d  dfma warps/smsp %d : %8.3f ms   DMMA.884 per clk per smsp %.4f (cycles each %.1f)   "
           "DFMA warp-instr per clk per smsp %.4f (cycles each %.1f)\n", c.wd, c.wf, ms, dmma_per_clk,
           dmma_per_clk ? 1 / dmma_per_clk : 0, dfma_per_clk, dfma_per_clk ? 1 / dfma_per_clk : 0);
  }
  printf("(cycles use the nominal clock %d kHz)\n", clk);
  return 0;
}
