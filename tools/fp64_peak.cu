// Measures the FP64 peaks the assembly kernel is judged against (MEASURED_PEAKS.json only has
// HBM + bf16): vector DFMA throughput and DMMA (mma.sync.m8n8k4.f64) throughput on this GPU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
  double x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k)
    x[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 16; ++k)
      x[k] = fma(x[k], a, b);
  double s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k)
    s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double *out, int iters, double a, double b)
{
  double c[8][2];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    c[k][0] = c[k][1] = 0.0;
  double av = a + threadIdx.x * 1e-6, bv = b;
  for (int i = 0; i < iters; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1])
                   : "d"(av), "d"(bv));
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *  out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024})
    for (int bps : {1, 2})
      {
        if (threads * bps > 2048)
          continue;
        const int grid = sms * bps, iters = 20000;
        float     ms;
        dfma_kernel<<<grid, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        dfma_kernel<<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 16 * double(iters) * grid * threads;
        printf("DFMA  threads/blk %4d blk/SM %d : %.2f TFLOP/s\n", threads, bps, fl / ms * 1e-9);
        dmma_kernel<<<grid, threads>>>(out, 100, 1.0000001, 1e-9);
        cudaEventRecord(e0);
        dmma_kernel<<<grid, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl2 = 2.0 * 8 * 8 * 4 * 8 * double(iters) * grid * (threads / 32);
        printf("DMMA  threads/blk %4d blk/SM %d : %.2f TFLOP/s\n", threads, bps, fl2 / ms * 1e-9);
      }
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  return 0;
}
