// Dependent-issue latencies of the FP64 instructions the window kernels chain (one warp, clock64):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu && tools/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b) {
  double x = a + threadIdx.x * 1e-9;
  __shared__ double sm[64];
  sm[threadIdx.x] = x;
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 64; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (OP == 0) x = fma(x, b, a);                                   // DFMA
      if (OP == 1) x = 1.0 / x + a;                                    // division + add
      if (OP == 2) x = sqrt(x) + a;                                    // sqrt + add
      if (OP == 3) x = rsqrt(x) + a;                                   // rsqrt + add
      if (OP == 4) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + a;   // 64-bit shuffle + add
      if (OP == 5) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31] + a; __syncwarp(); }  // STS+LDS
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* d; long long* c; cudaMalloc(&d, 256); cudaMalloc(&c, 8);
  const char* names[] = {"DFMA", "div+DADD", "sqrt+DADD", "rsqrt+DADD", "SHFL64+DADD", "STS+sync+LDS+DADD+sync"};
  for (int op = 0; op < 6; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: chain<0><<<1, 32>>>(d, c, 1.0000001, 0.9999999); break;
        case 1: chain<1><<<1, 32>>>(d, c, 1.0000001, 0.9999999); break;
        case 2: chain<2><<<1, 32>>>(d, c, 1.0000001, 0.9999999); break;
        case 3: chain<3><<<1, 32>>>(d, c, 1.0000001, 0.9999999); break;
        case 4: chain<4><<<1, 32>>>(d, c, 1.0000001, 0.9999999); break;
        case 5: chain<5><<<1, 32>>>(d, c, 1.0000001, 0.9999999); break;
      }
      cudaDeviceSynchronize();
    }
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"cycles_per_dependent_op\": %.1f}\n", names[op], (double)h / 1024.0);
  }
  return 0;
}
