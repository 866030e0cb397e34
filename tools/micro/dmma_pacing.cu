// dev tool: can a DMMA warp be paced so that a DFMA warp on the same sub-partition gets a fixed share of the FP64 datapath?
// One DMMA warp + one DFMA warp per sub-partition (8 warps per CTA); the DMMA warp inserts a pacing instruction after every DMMA (KIND) or every
// EVERY-th DMMA.  Reports the datapath share of each role over a fixed clock window.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dmma_pacing tools/micro/dmma_pacing.cu && build/dmma_pacing
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__device__ __forceinline__ void pace(unsigned& x)
{
  if (KIND == 1) __syncwarp();
  if (KIND == 2) asm volatile("nanosleep.u32 0;");
  if (KIND == 3) asm volatile("mov.u32 %0, %%clock;" : "=r"(x));
  if (KIND == 4) asm volatile("{.reg .pred p; add.u32 %0, %0, 1; add.u32 %0, %0, 1; setp.eq.u32 p, %0, 0x7fffffff; @p trap;}" : "+r"(x));
  if (KIND == 5) asm volatile("bar.warp.sync 0xffffffff; bar.warp.sync 0xffffffff;");
  if (KIND == 6) asm volatile("{.reg .b32 t; shfl.sync.idx.b32 t, %0, 0, 31, 0xffffffff; add.u32 %0, %0, t;}" : "+r"(x));
}

template <int KIND, int EVERY, int DFMA_ILP>
__global__ void share(double* out, long long window, unsigned long long* done)
{
  const int warp = threadIdx.x >> 5;
  double acc = 0;
  unsigned long long n = 0;
  unsigned x = threadIdx.x;
  const long long t0 = clock64();
  if (warp < 4)
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
        {
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[k][0]), "+d"(d[k][1]) : "d"(a), "d"(b));
          if ((r * 8 + k) % EVERY == EVERY - 1) pace<KIND>(x);
        }
      n += 32;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc += d[k][0] + d[k][1];
  }
  else
  {
    double v[DFMA_ILP];
#pragma unroll
    for (int k = 0; k < DFMA_ILP; k++) v[k] = threadIdx.x + k;
    const double a = 1.0000001, b = 1e-9;
    while (clock64() - t0 < window)
    {
#pragma unroll
      for (int r = 0; r < 64 / DFMA_ILP; r++)
#pragma unroll
        for (int k = 0; k < DFMA_ILP; k++) v[k] = fma(v[k], a, b);
      n += 64;
    }
#pragma unroll
    for (int k = 0; k < DFMA_ILP; k++) acc += v[k];
  }
  out[threadIdx.x] = acc + x;
  if ((threadIdx.x & 31) == 0) done[warp] = n;
}

template <int KIND, int EVERY, int ILP>
void run(double* out, unsigned long long* done, const char* name)
{
  const long long window = 2000000;
  unsigned long long h[8];
  share<KIND, EVERY, ILP><<<1, 256>>>(out, window, done);
  cudaMemcpy(h, done, 64, cudaMemcpyDeviceToHost);
  printf("%-34s every %d DMMA, DFMA ILP %d: DMMA %5.1f%%  DFMA %5.1f%%  sum %5.1f%%\n", name, EVERY, ILP, 100.0 * h[0] * 16 / window, 100.0 * h[4] * 2 / window,
         100.0 * (h[0] * 16 + h[4] * 2) / window);
}

int main()
{
  double* out;
  unsigned long long* done;
  cudaMalloc(&out, 8 * 4096);
  cudaMalloc(&done, 64);
#define ALL(ILP)                                          \
  run<0, 1, ILP>(out, done, "no pacing");                 \
  run<1, 1, ILP>(out, done, "__syncwarp");                \
  run<5, 1, ILP>(out, done, "2 x bar.warp.sync");         \
  run<3, 1, ILP>(out, done, "read %clock");               \
  run<4, 1, ILP>(out, done, "2 adds + setp + @p trap");   \
  run<6, 1, ILP>(out, done, "shfl + add");                \
  run<2, 1, ILP>(out, done, "nanosleep 0");               \
  run<2, 4, ILP>(out, done, "nanosleep 0");               \
  run<2, 8, ILP>(out, done, "nanosleep 0");               \
  run<6, 2, ILP>(out, done, "shfl + add");
  ALL(8)
  ALL(2)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
