// dev tool: dependent-issue latency and single-warp throughput of DFMA / DMMA on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dfma_latency tools/micro/dfma_latency.cu && build/dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, int iters, long long* cyc)
{
  double x[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) x[k] = threadIdx.x + k;
  const double a = 1.0000001, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
__global__ void dmma_chain(double* out, int iters, long long* cyc)
{
  double d[ILP][2];
#pragma unroll
  for (int k = 0; k < ILP; k++) d[k][0] = d[k][1] = 0.0;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++)
  {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < ILP; k++)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[k][0]), "+d"(d[k][1]) : "d"(a), "d"(b));
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s += d[k][0] + d[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP>
void run(double* out, long long* cyc, int warps)
{
  const int iters = 2000;
  long long h;
  dfma_chain<ILP><<<1, 32 * warps>>>(out, iters, cyc);
  dfma_chain<ILP><<<1, 32 * warps>>>(out, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA warps/SM %2d ILP %d: %.2f cycles per DFMA per warp  (%.2f cycles between issues on a sub-partition)\n", warps, ILP, (double)h / (iters * 8.0 * ILP),
         (double)h / (iters * 8.0 * ILP) / ((warps + 3) / 4));
  dmma_chain<ILP><<<1, 32 * warps>>>(out, iters, cyc);
  dmma_chain<ILP><<<1, 32 * warps>>>(out, iters, cyc);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA warps/SM %2d ILP %d: %.2f cycles per DMMA per warp\n", warps, ILP, (double)h / (iters * 4.0 * ILP));
}

int main()
{
  double* out;
  long long* cyc;
  cudaMalloc(&out, 8 * 1024);
  cudaMalloc(&cyc, 8);
  for (int warps : {1, 4, 8})
  {
    run<1>(out, cyc, warps);
    run<2>(out, cyc, warps);
    run<4>(out, cyc, warps);
    run<8>(out, cyc, warps);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
