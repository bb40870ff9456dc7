// tools/pcie_bench.cu -- PCIe ceilings for the host-operand (e2e) path: contiguous vs strided (2-D) copies, one
// direction and both directions at once.  Build: nvcc -O3 -o gpurun_out/pcie_bench tools/pcie_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
static double ms(cudaEvent_t a, cudaEvent_t b) { float t; cudaEventElapsedTime(&t, a, b); return t; }
int main() {
  const size_t rows = 4u << 20, pitch = 512, total = rows * pitch;      // 2 GiB host matrix, 512-byte rows
  char *h, *h2, *d, *d2;
  CK(cudaHostAlloc(&h, total, cudaHostAllocDefault));
  CK(cudaHostAlloc(&h2, total, cudaHostAllocDefault));
  CK(cudaMalloc(&d, total)); CK(cudaMalloc(&d2, total));
  cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a, s1); CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s1)); cudaEventRecord(b, s1); CK(cudaEventSynchronize(b));
    if (rep) printf("H2D contiguous 2 GiB        : %6.1f GB/s\n", total / ms(a, b) / 1e6);
    cudaEventRecord(a, s1); CK(cudaMemcpyAsync(h2, d2, total, cudaMemcpyDeviceToHost, s1)); cudaEventRecord(b, s1); CK(cudaEventSynchronize(b));
    if (rep) printf("D2H contiguous 2 GiB        : %6.1f GB/s\n", total / ms(a, b) / 1e6);
    cudaEventRecord(a, s1);
    CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpyAsync(h2, d2, total, cudaMemcpyDeviceToHost, s2));
    cudaEventRecord(b, s1); CK(cudaStreamSynchronize(s2)); CK(cudaEventSynchronize(b));
    CK(cudaDeviceSynchronize());
    if (rep) printf("H2D + D2H concurrently      : see next line\n");
    cudaEvent_t c; cudaEventCreate(&c);
    cudaEventRecord(a, s1); cudaStreamWaitEvent(s2, a, 0);
    CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s1));
    CK(cudaMemcpyAsync(h2, d2, total, cudaMemcpyDeviceToHost, s2));
    cudaEventRecord(c, s2); cudaStreamWaitEvent(s1, c, 0); cudaEventRecord(b, s1); CK(cudaEventSynchronize(b));
    if (rep) printf("  both directions, 2+2 GiB  : %6.1f GB/s aggregate (%.1f ms)\n", 2.0 * total / ms(a, b) / 1e6, ms(a, b));
    for (size_t w : {64, 128, 256}) {
      cudaEventRecord(a, s1); CK(cudaMemcpy2DAsync(d, w, h, pitch, w, rows, cudaMemcpyHostToDevice, s1)); cudaEventRecord(b, s1); CK(cudaEventSynchronize(b));
      if (rep) printf("H2D 2-D width %3zu B of 512 B : %6.1f GB/s\n", w, rows * w / ms(a, b) / 1e6);
      cudaEventRecord(a, s1); CK(cudaMemcpy2DAsync(h2, pitch, d2, w, w, rows, cudaMemcpyDeviceToHost, s1)); cudaEventRecord(b, s1); CK(cudaEventSynchronize(b));
      if (rep) printf("D2H 2-D width %3zu B of 512 B : %6.1f GB/s\n", w, rows * w / ms(a, b) / 1e6);
    }
  }
  return 0;
}
