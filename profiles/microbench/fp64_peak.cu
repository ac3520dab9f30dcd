// Microbenchmark: sustained FP64 throughput on sm_100a via (a) DFMA register chains and
// (b) mma.sync.m8n8k4.f64 (DMMA).  Used to decide how the Schur SYRK kernel should issue its math.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
  double c[NT][2];
#pragma unroll
  for (int i = 0; i < NT; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NT; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32, blocks = sms;
    float ms = time_it([&] { dfma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 16 * iters * (double)threads * blocks;
    printf("DFMA  16 acc/thread  %2d warps/SM: %7.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    ms = time_it([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 256 * 8 * iters * (double)warps * blocks;
    printf("DMMA   8 tiles/warp  %2d warps/SM: %7.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
    ms = time_it([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 256 * 16 * iters * (double)warps * blocks;
    printf("DMMA  16 tiles/warp  %2d warps/SM: %7.3f ms  %7.2f TFLOP/s\n", warps, ms, fl / ms * 1e-9);
  }
  return 0;
}
