// dev tool: does a DFMA whose upper half-warp is inactive occupy the FP64 datapath of a sub-partition for one pass (16 lanes) or two?
// and how does the warp scheduler share the datapath between DMMA warps and DFMA warps on one sub-partition?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dfma_halfwarp tools/micro/dfma_halfwarp.cu && build/dfma_halfwarp
#include <cstdio>
#include <cuda_runtime.h>

// every warp runs `iters` x 64 independent-enough DFMAs (ILP 8); lanes >= active leave before the loop
__global__ void dfma_lanes(double* out, int iters, int active, long long* cyc)
{
  const int lane = threadIdx.x & 31;
  if (lane >= active) return;
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
  const double a = 1.0000001, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// warps [0, n_mma) issue DMMAs, warps [n_mma, n_mma + n_dfma) issue DFMAs (ILP 8); all on sub-partition (warp % 4); each role reports the
// work it completed in a fixed time window (clock64 deadline), i.e. its share of the datapath
__global__ void share(double* out, long long window, int n_mma, unsigned long long* done, int n_dfma_first = 0)
{
  // n_dfma_first > 0: the DFMA warps take the LOWEST warp ids instead (does the scheduler's choice depend on the warp id?)
  const int warp = threadIdx.x >> 5;
  double acc = 0;
  unsigned long long n = 0;
  const long long t0 = clock64();
  if (n_dfma_first > 0 ? warp >= n_dfma_first : warp < n_mma)
  {
    double d[8][2];
#pragma unroll
    for (int k = 0; k < 8; k++) d[k][0] = d[k][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    while (clock64() - t0 < window)
    {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int k = 0; k < 8; k++)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[k][0]), "+d"(d[k][1]) : "d"(a), "d"(b));
      n += 32;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc += d[k][0] + d[k][1];
  }
  else
  {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x + k;
    const double a = 1.0000001, b = 1e-9;
    while (clock64() - t0 < window)
    {
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int k = 0; k < 8; k++) x[k] = fma(x[k], a, b);
      n += 64;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc += x[k];
  }
  out[threadIdx.x] = acc;
  if ((threadIdx.x & 31) == 0) done[warp] = n;
}

int main()
{
  double* out;
  long long* cyc;
  unsigned long long* done;
  cudaMalloc(&out, 8 * 4096);
  cudaMalloc(&cyc, 8);
  cudaMalloc(&done, 8 * 64);
  const int iters = 4000;
  for (int warps : {1, 4, 8})
    for (int active : {32, 16, 8})
    {
      long long h;
      dfma_lanes<<<1, 32 * warps>>>(out, iters, active, cyc);
      dfma_lanes<<<1, 32 * warps>>>(out, iters, active, cyc);
      cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("DFMA %d warps/SM, %2d active lanes: %.2f cycles per DFMA per warp\n", warps, active, (double)h / (iters * 64.0));
    }
  // datapath sharing on ONE sub-partition: warps 0,4,8,.. share sub-partition 0 when the block is launched with warps only there -> use 4*k warps and
  // look at sub-partition 0 (warps 0, 4, 8, ...)
  const long long window = 2000000;
  for (int n_mma_per : {1, 2})
    for (int n_dfma_per : {1, 2, 3})
    {
      const int n_mma = 4 * n_mma_per, n_dfma = 4 * n_dfma_per;
      unsigned long long h[64];
      share<<<1, 32 * (n_mma + n_dfma)>>>(out, window, n_mma, done);
      cudaMemcpy(h, done, 8 * (n_mma + n_dfma), cudaMemcpyDeviceToHost);
      unsigned long long m = 0, f = 0;
      for (int w = 0; w < n_mma + n_dfma; w += 4) (w < n_mma ? m : f) += h[w];
      printf("sub-partition 0 with %d DMMA warp(s) + %d DFMA warp(s): DMMA %.1f%% of the datapath (16 cyc each), DFMA %.1f%% (2 cyc each)\n", n_mma_per,
             n_dfma_per, 100.0 * m * 16 / window, 100.0 * f * 2 / window);
    }
  for (int n_mma_per : {1, 2})
    for (int n_dfma_per : {1, 2})
    {
      const int n_mma = 4 * n_mma_per, n_dfma = 4 * n_dfma_per;
      unsigned long long h[64];
      share<<<1, 32 * (n_mma + n_dfma)>>>(out, window, n_mma, done, n_dfma);
      cudaMemcpy(h, done, 8 * (n_mma + n_dfma), cudaMemcpyDeviceToHost);
      unsigned long long m = 0, f = 0;
      for (int w = 0; w < n_mma + n_dfma; w += 4) (w >= n_dfma ? m : f) += h[w];
      printf("DFMA warps FIRST: sub-partition 0 with %d DMMA warp(s) + %d DFMA warp(s): DMMA %.1f%% of the datapath, DFMA %.1f%%\n", n_mma_per, n_dfma_per,
             100.0 * m * 16 / window, 100.0 * f * 2 / window);
    }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
