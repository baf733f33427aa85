// Measures the FP64 denominators MEASURED_PEAKS.json lacks: sustained DFMA
// and DMMA (mma.sync.m8n8k4.f64) throughput of the whole GPU, with CUDA events.
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int j = 0; j < 8; ++j) { c[j][0] = threadIdx.x; c[j][1] = j; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int tpb : {256, 512, 1024}) {
    int blocks = sms * (2048 / tpb);
    int iters = 20000;
    float ms;
    dfma_kernel<<<blocks, tpb>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, tpb>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 8 * iters * (double)blocks * tpb;
    printf("{\"kernel\":\"dfma\",\"tpb\":%d,\"blocks\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", tpb, blocks, ms, flops / ms * 1e-9);
    dmma_kernel<<<blocks, tpb>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, tpb>>>(out, iters / 4, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    flops = 2.0 * 8 * 8 * 4 * 8 * (iters / 4) * (double)blocks * (tpb / 32);
    printf("{\"kernel\":\"dmma884\",\"tpb\":%d,\"blocks\":%d,\"ms\":%.3f,\"tflops\":%.2f}\n", tpb, blocks, ms, flops / ms * 1e-9);
  }
  printf("{\"sms\":%d,\"clock_khz\":%d,\"err\":\"%s\"}\n", sms, p.clockRate, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
