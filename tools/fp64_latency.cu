// Micro-benchmark (one warp): dependent-issue latency and single-warp throughput of the fp64
// operations on the DDP critical path.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

template<int CHAINS>
__global__ void dfma_kernel(double * out, long long * cyc, double a, double b, int iters)
{
  double v[CHAINS];
  for(int c = 0; c < CHAINS; c++) v[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
#pragma unroll
    for(int c = 0; c < CHAINS; c++) v[c] = fma(v[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for(int c = 0; c < CHAINS; c++) s += v[c];
  out[threadIdx.x] = s;
  if(threadIdx.x == 0) *cyc = t1 - t0;
}

template<int OP>
__global__ void op_kernel(double * out, long long * cyc, double a, int iters)
{
  double v = 0.3 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
    if(OP == 0) v = 1.0 / (v + a);
    if(OP == 1) v = sqrt(v + a);
    if(OP == 2)
    {
      double s, c;
      sincos(v, &s, &c);
      v = s + c * a;
    }
    if(OP == 3) v = sin(v) + a;
    if(OP == 4) v = __drcp_rn(v + a);
    if(OP == 5) v = v * a + 0.25; // DMUL+DADD or DFMA
    if(OP == 6) v = __shfl_xor_sync(0xffffffffu, v, 1) + a;
    if(OP == 7) v = rsqrt(v + a);
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if(threadIdx.x == 0) *cyc = t1 - t0;
}


// DFMA issue rate of ONE warp with only `active` lanes enabled (does the fp64 pipe skip an idle half-warp?)
__global__ void dfma_partial_kernel(double * out, long long * cyc, double a, double b, int iters, int active)
{
  if((int)threadIdx.x >= active) return;
  double v[16];
  for(int c = 0; c < 16; c++) v[c] = threadIdx.x * 1e-3 + c;
  long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
#pragma unroll
    for(int c = 0; c < 16; c++) v[c] = fma(v[c], a, b);
  }
  long long t1 = clock64();
  double s = 0;
  for(int c = 0; c < 16; c++) s += v[c];
  out[threadIdx.x] = s;
  if(threadIdx.x == 0) *cyc = t1 - t0;
}

// Round trip of one value through shared memory between the 4 warps of a CTA: STS, bar.sync, LDS of the
// neighbour warp's value, one dependent DFMA (the exchange step of a column-split Riccati sweep).
__global__ void exchange_kernel(double * out, long long * cyc, double a, int iters)
{
  __shared__ double buf[4][32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  double v = 0.3 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
    buf[w][l] = v;
    __syncthreads();
    v = fma(buf[(w + 1) & 3][l], a, 0.25);
    __syncthreads();
  }
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if(threadIdx.x == 0) *cyc = t1 - t0;
}

int main()
{
  double * out;
  long long * cyc;
  cudaMalloc(&out, 1024 * 8);
  cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
#define RUN_DFMA(C)                                                                 \
  dfma_kernel<C><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters);                           \
  cudaDeviceSynchronize();                                                          \
  dfma_kernel<C><<<1, 32>>>(out, cyc, 0.999, 1e-3, iters);                           \
  cudaDeviceSynchronize();                                                          \
  printf("DFMA chains=%d: %.2f cyc/iter, %.2f cyc/DFMA\n", C, double(*cyc) / iters, double(*cyc) / iters / C);
  RUN_DFMA(1)
  RUN_DFMA(2)
  RUN_DFMA(4)
  RUN_DFMA(8)
  RUN_DFMA(16)
  const char * names[] = {"div 1/x", "sqrt", "sincos", "sin", "drcp_rn", "fma", "shfl f64", "rsqrt"};
#define RUN_OP(O)                                        \
  op_kernel<O><<<1, 32>>>(out, cyc, 0.5, iters);          \
  cudaDeviceSynchronize();                               \
  op_kernel<O><<<1, 32>>>(out, cyc, 0.5, iters);          \
  cudaDeviceSynchronize();                               \
  printf("%-10s: %.1f cyc/op (dependent)\n", names[O], double(*cyc) / iters);
  RUN_OP(0)
  RUN_OP(1)
  RUN_OP(2)
  RUN_OP(3)
  RUN_OP(4)
  RUN_OP(5)
  RUN_OP(6)
  RUN_OP(7)
  // 4 warps on one SM (one per scheduler) and 8 warps: aggregate DFMA throughput
  for(int w : {4, 8, 16})
  {
    dfma_kernel<8><<<1, 32 * w>>>(out, cyc, 0.999, 1e-3, iters);
    cudaDeviceSynchronize();
    printf("DFMA 8 chains x %d warps: %.2f cyc/iter (warp 0) => %.2f warp-DFMA/cyc/SM\n", w, double(*cyc) / iters,
           8.0 * w / (double(*cyc) / iters));
  }
  for(int act : {32, 16, 8, 4})
  {
    dfma_partial_kernel<<<1, 32>>>(out, cyc, 0.999, 1e-3, iters, act);
    cudaDeviceSynchronize();
    dfma_partial_kernel<<<1, 32>>>(out, cyc, 0.999, 1e-3, iters, act);
    cudaDeviceSynchronize();
    printf("DFMA 16 chains, %2d active lanes: %.2f cyc/DFMA\n", act, double(*cyc) / iters / 16);
  }
  exchange_kernel<<<1, 128>>>(out, cyc, 0.999, iters);
  cudaDeviceSynchronize();
  exchange_kernel<<<1, 128>>>(out, cyc, 0.999, iters);
  cudaDeviceSynchronize();
  printf("smem exchange (STS, bar.sync, LDS, DFMA, bar.sync) x 4 warps: %.1f cyc/round\n", double(*cyc) / iters);
  return 0;
}
