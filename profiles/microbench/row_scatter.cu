// Microbenchmark: write bandwidth of 160-byte (5-sector) rows with STG.256 as a function of how
// rows are ordered across a warp.  Models resjac_kernel's Jacobian-row store: lane l of a warp
// writes row perm[q] (32-byte aligned, full sectors).  run = number of consecutive rows written by
// consecutive lanes before jumping elsewhere.
#include <cstdio>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__device__ __forceinline__ void st256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
template <int SECT>
__global__ void scatter_rows(const int* __restrict__ perm, int n, double* __restrict__ out, int stride_d) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n; q += gridDim.x * blockDim.x) {
    double* dst = out + (size_t)perm[q] * stride_d;
    double v = q;
#pragma unroll
    for (int s = 0; s < SECT; ++s) st256(dst + 4 * s, v, v + 1, v + 2, v + 3);
  }
}

int main() {
  const int n = 2000000;
  std::mt19937 rng(1);
  int* d_perm; cudaMalloc(&d_perm, n * sizeof(int));
  double* out; cudaMalloc(&out, (size_t)n * 24 * 8);
  double* flush; cudaMalloc(&flush, 256u << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int run : {1, 2, 4, 8, 16, 32, 0}) {
    std::vector<int> perm(n);
    if (run == 0) { for (int i = 0; i < n; ++i) perm[i] = i; }
    else {
      int nb = n / run; std::vector<int> blocks(nb);
      for (int i = 0; i < nb; ++i) blocks[i] = i;
      std::shuffle(blocks.begin(), blocks.end(), rng);
      for (int i = 0; i < nb; ++i) for (int k = 0; k < run; ++k) perm[i * run + k] = blocks[i] * run + k;
    }
    cudaMemcpy(d_perm, perm.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant) {
      int stride = variant == 0 ? 20 : 24;
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(flush, rep, 256u << 20);
        cudaEventRecord(e0);
        if (variant == 0) scatter_rows<5><<<148 * 8, 256>>>(d_perm, n, out, stride);
        else scatter_rows<6><<<148 * 8, 256>>>(d_perm, n, out, stride);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
      }
      double bytes = (double)n * stride * 8;
      printf("run %2d rows/%s: %7.1f us  %7.1f GB/s payload\n", run, variant == 0 ? "160B" : "192B(64B-aligned)", best * 1e3, bytes / best * 1e-6);
    }
  }
  return 0;
}
