// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/probe/tma_gather tools/probe/tma_gather.cu
// Probe: random 8-byte gathers from an L2-resident 8 MB table -- LSU loads against 16-byte bulk copies (TMA path) and a mix.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int TILE = 2048, THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// tma_share: of every 8 gathers a thread owns, how many go through cp.async.bulk (0..8)
template <int TMA_OF_8>
__global__ void __launch_bounds__(THREADS, 6) gather_kernel(const double* __restrict__ x, const int* __restrict__ idx, int n, double* __restrict__ out) {
  __shared__ __align__(16) double2 stage[TMA_OF_8 > 0 ? TILE / 8 * TMA_OF_8 : 1];
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * TILE;
  if (TMA_OF_8 > 0) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)(TILE / 8 * TMA_OF_8 * 16)) : "memory");
    __syncthreads();
  }
  int c[8];
#pragma unroll
  for (int u = 0; u < 8; u++) c[u] = idx[base + u * THREADS + tid];
  double s = 0;
#pragma unroll
  for (int u = 0; u < TMA_OF_8; u++) {
    const char* src = (const char*)x + (((size_t)c[u] * 8) & ~(size_t)15);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 16, [%2];" ::"r"(smem_u32(&stage[u * THREADS + tid])),
                 "l"(src), "r"(smem_u32(&bar)) : "memory");
  }
#pragma unroll
  for (int u = TMA_OF_8; u < 8; u++) s += __ldg(x + c[u]);
  if (TMA_OF_8 > 0) {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
#pragma unroll
    for (int u = 0; u < TMA_OF_8; u++) {
      const double2 v = stage[u * THREADS + tid];
      s += (c[u] & 1) ? v.y : v.x;
    }
  }
  out[base / TILE * THREADS + tid] = s;
}

template <int T8> static float run(const double* x, const int* idx, int n, double* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    gather_kernel<T8><<<n / TILE, THREADS>>>(x, idx, n, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const int m = 1 << 20, n = TILE * 4883;      // 10 M gathers from a 1 Mi-entry fp64 table
  std::vector<int> h(n); std::vector<double> hx(m);
  uint64_t st = 12345;
  for (int i = 0; i < n; i++) { st = st * 6364136223846793005ull + 1442695040888963407ull; h[i] = (int)((st >> 33) % m); }
  for (int i = 0; i < m; i++) hx[i] = i;
  double *x, *out; int* idx;
  CK(cudaMalloc(&x, m * 8)); CK(cudaMalloc(&idx, (size_t)n * 4)); CK(cudaMalloc(&out, (size_t)n / 8 * 8));
  CK(cudaMemcpy(x, hx.data(), m * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(idx, h.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  float t;
  t = run<0>(x, idx, n, out); printf("LSU only          %.1f us  %.0f G gathers/s\n", t * 1e3, n / t / 1e6);
  std::vector<double> ref(n / 8), got(n / 8);
  CK(cudaMemcpy(ref.data(), out, (size_t)n / 8 * 8, cudaMemcpyDeviceToHost));
#define RUN(T8) t = run<T8>(x, idx, n, out); CK(cudaDeviceSynchronize()); CK(cudaMemcpy(got.data(), out, (size_t)n / 8 * 8, cudaMemcpyDeviceToHost)); \
  printf("bulk %d of 8       %.1f us  %.0f G gathers/s  %s\n", T8, t * 1e3, n / t / 1e6, got == ref ? "same sums" : "DIFFERENT");
  RUN(1) RUN(2) RUN(3) RUN(4) RUN(8)
  return 0;
}
