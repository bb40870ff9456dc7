// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/probe/pcie2d tools/probe/pcie2d.cu
// PCIe probe: strided (2D) pinned<->device copies of a column slice of a row-major [rows x 128] fp32 matrix, alone and duplex.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const size_t rows = 4u << 20, pitch = 512;
  char *h, *h2, *d, *d2;
  CK(cudaMallocHost(&h, rows * pitch)); CK(cudaMallocHost(&h2, rows * pitch));
  CK(cudaMalloc(&d, rows * pitch)); CK(cudaMalloc(&d2, rows * pitch));
  cudaStream_t up, dn; CK(cudaStreamCreate(&up)); CK(cudaStreamCreate(&dn));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (size_t w : {512, 256, 128, 64}) {
    for (int mode = 0; mode < 3; mode++) {     // 0: H2D, 1: D2H, 2: both at once
      float best = 1e9;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, up));
        CK(cudaStreamWaitEvent(dn, a, 0));
        if (mode != 1) CK(cudaMemcpy2DAsync(d, w, h, pitch, w, rows, cudaMemcpyHostToDevice, up));
        if (mode != 0) CK(cudaMemcpy2DAsync(h2, pitch, d2, w, w, rows, cudaMemcpyDeviceToHost, dn));
        CK(cudaEventRecord(b, dn)); CK(cudaStreamWaitEvent(up, b, 0));
        CK(cudaEventRecord(b, up));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
      }
      printf("width %zu B  %s  %.2f ms  %.1f GB/s per direction\n", w, mode == 0 ? "H2D" : mode == 1 ? "D2H" : "duplex", best, rows * w / best / 1e6);
    }
  }
  return 0;
}
