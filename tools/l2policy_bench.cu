// tools/l2policy_bench.cu -- do L2 eviction-priority hints protect a hot gather table from a concurrent stream?
// A 48 MiB table is gathered at random (512-byte rows) while a 2 GiB array streams through; L2 is 126 MB.
// Reports the time of the mixed kernel with (a) no hints, (b) table evict_last + stream evict_first.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/l2policy tools/l2policy_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <bool HINT>
__global__ void __launch_bounds__(256) mixed(const float4* __restrict__ table, const int* __restrict__ idx, long long nidx,
                                             const float4* __restrict__ strm, size_t nstream, float* out) {
  uint64_t keep, first;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(first));
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float acc = 0;
  const long long per = (nidx + nwarps - 1) / nwarps;
  const size_t sper = (nstream + nwarps - 1) / nwarps;
  size_t sp = warp * sper;
  for (long long p = warp * per; p < min(nidx, (warp + 1) * per); p += 32) {
    int my = (p + lane < nidx) ? idx[p + lane] : 0;
    for (int j = 0; j < 32; j += 4) {
      float4 v[4], s;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        int r = __shfl_sync(0xffffffffu, my, j + u);
        const float4* a = table + (size_t)r * 32 + lane;
        if (HINT) asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(a), "l"(keep));
        else v[u] = __ldg(a);
      }
      // one streaming 512-byte load per 4 gathers (ratio chosen so that the stream is ~the size of the gathers)
      const float4* b = strm + (sp % nstream) + lane;
      sp += 32;
      if (HINT) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w) : "l"(b), "l"(first));
      else asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w) : "l"(b));
#pragma unroll
      for (int u = 0; u < 4; u++) acc += v[u].x + v[u].w;
      acc += s.x;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const size_t table_bytes = (size_t)96 << 20, stream_bytes = (size_t)2 << 30;
  const long long nidx = 16 << 20;
  float4 *table, *strm; int* idx; float* out;
  CK(cudaMalloc(&table, table_bytes)); CK(cudaMalloc(&strm, stream_bytes)); CK(cudaMalloc(&idx, nidx * 4)); CK(cudaMalloc(&out, 64));
  CK(cudaMemset(table, 0, table_bytes)); CK(cudaMemset(strm, 0, stream_bytes));
  std::vector<int> h(nidx);
  uint64_t s = 88172645463325252ull;
  for (long long i = 0; i < nidx; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (table_bytes / 512)); }
  CK(cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; rep++) {
    float t0, t1;
    cudaEventRecord(a); mixed<false><<<148 * 8, 256>>>(table, idx, nidx, strm, stream_bytes / 16, out); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&t0, a, b);
    cudaEventRecord(a); mixed<true><<<148 * 8, 256>>>(table, idx, nidx, strm, stream_bytes / 16, out); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    cudaEventElapsedTime(&t1, a, b);
    printf("rep %d: 96 MiB table gathers (8 GiB) + 2 GiB stream: no hints %.3f ms, evict_last/evict_first %.3f ms\n", rep, t0, t1);
  }
  return 0;
}
