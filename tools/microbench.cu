// tools/microbench.cu -- machine ceilings for the access patterns of the hot path (run under gpurun):
//   stream copy / read, random ROW gathers (SpMM / SDDMM / MTTKRP pattern) for several row sizes, table sizes
//   (L2-resident ... HBM-resident) and loads in flight, and random 8-byte gathers (SpMV x pattern).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void read_kernel(const float4* __restrict__ a, float* out, size_t n) {
  float s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = a[i];
    s += v.x + v.y + v.z + v.w;
  }
  if (s == 123.456f) out[0] = s;
}

// each warp gathers `per_warp` rows of ROWB bytes; lanes cover the row with 16-byte loads; U rows in flight
template <int ROWB, int U>
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ table, const int* __restrict__ idx,
                                                          long long nidx, int per_warp, float* out) {
  constexpr int LPR = ROWB / 16;            // lanes per row
  constexpr int RPW = 32 / LPR;             // rows per warp instruction
  const int lane = threadIdx.x & 31;
  const long long w = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  long long p0 = w * per_warp;
  if (p0 >= nidx) return;
  float acc = 0;
  const int sub = lane / LPR, off = lane % LPR;
  for (int b = 0; b < per_warp; b += 32) {
    int my = (p0 + b + lane < nidx) ? __ldg(idx + p0 + b + lane) : 0;
    for (int j = 0; j < 32; j += U * RPW) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        int r = __shfl_sync(0xffffffffu, my, j + u * RPW + sub);
        v[u] = __ldg(table + (size_t)r * LPR + off);
      }
#pragma unroll
      for (int u = 0; u < U; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

__global__ void gather8_kernel(const double* __restrict__ x, const int* __restrict__ idx, long long n, double* out) {
  double s = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += __ldg(x + __ldg(idx + i));
  if (s == 123.456) out[0] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

template <int ROWB, int U>
static void run_gather(const float4* table, size_t table_bytes, const int* idx, long long nidx, float* out, const char* tag) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int per_warp = 64;
  long long warps = (nidx + per_warp - 1) / per_warp;
  int grid = (int)((warps * 32 + 255) / 256);
  gather_rows_kernel<ROWB, U><<<grid, 256>>>(table, idx, nidx, per_warp, out);
  float best = 1e9;
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(a);
    gather_rows_kernel<ROWB, U><<<grid, 256>>>(table, idx, nidx, per_warp, out);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    best = fminf(best, time_ms(a, b));
  }
  printf("gather rows  %-8s row=%4dB U=%2d table=%7.0f MB : %8.1f GB/s  (%6.1f Mrows/s, %.3f ms)\n", tag, ROWB, U,
         table_bytes / 1e6, nidx * (double)ROWB / best / 1e6, nidx / best / 1e3, best);
}

int main() {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  float* out; CK(cudaMalloc(&out, 64));
  // ---- stream ------------------------------------------------------------------------------------------------
  size_t n = (size_t)1 << 30;   // 1 GiB each
  float4 *x, *y;
  CK(cudaMalloc(&x, n)); CK(cudaMalloc(&y, n));
  CK(cudaMemset(x, 1, n)); CK(cudaMemset(y, 0, n));
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(a); copy_kernel<<<148 * 16, 256>>>(x, y, n / 16); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    if (it == 2) printf("stream copy 1 GiB: %.1f GB/s (read+write)\n", 2.0 * n / time_ms(a, b) / 1e6);
  }
  for (int it = 0; it < 3; it++) {
    cudaEventRecord(a); read_kernel<<<148 * 16, 256>>>(x, out, n / 16); cudaEventRecord(b); CK(cudaEventSynchronize(b));
    if (it == 2) printf("stream read 1 GiB: %.1f GB/s\n", 1.0 * n / time_ms(a, b) / 1e6);
  }
  {
    size_t m = (size_t)64 << 20;   // L2-resident read
    read_kernel<<<148 * 16, 256>>>(x, out, m / 16);
    cudaEventRecord(a);
    for (int r = 0; r < 20; r++) read_kernel<<<148 * 16, 256>>>(x, out, m / 16);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    printf("L2-resident read 64 MiB x20: %.1f GB/s\n", 20.0 * m / time_ms(a, b) / 1e6);
  }
  // ---- row gathers -------------------------------------------------------------------------------------------
  const long long nidx = 32 << 20;          // 32 Mi rows gathered
  std::vector<int> h(nidx);
  int* idx; CK(cudaMalloc(&idx, nidx * 4));
  size_t tables[] = {(size_t)32 << 20, (size_t)96 << 20, (size_t)512 << 20, (size_t)2048 << 20};
  float4* table; CK(cudaMalloc(&table, tables[3]));
  CK(cudaMemset(table, 0, tables[3]));
  for (size_t tb : tables) {
    for (int rowb : {128, 256, 512}) {
      uint64_t s = 88172645463325252ull;
      long long nrows = tb / rowb;
      for (long long i = 0; i < nidx; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)nrows); }
      CK(cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice));
      if (rowb == 128) { run_gather<128, 2>(table, tb, idx, nidx, out, "uniform"); run_gather<128, 4>(table, tb, idx, nidx, out, "uniform"); }
      if (rowb == 256) { run_gather<256, 4>(table, tb, idx, nidx, out, "uniform"); run_gather<256, 8>(table, tb, idx, nidx, out, "uniform"); }
      if (rowb == 512) { run_gather<512, 4>(table, tb, idx, nidx, out, "uniform"); run_gather<512, 8>(table, tb, idx, nidx, out, "uniform");
                         run_gather<512, 16>(table, tb, idx, nidx, out, "uniform"); }
    }
  }
  // ---- 8-byte gathers from an 8 MB vector (SpMV x) -------------------------------------------------------------
  {
    uint64_t s = 1234567;
    for (long long i = 0; i < nidx; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % 1000000ull); }
    CK(cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice));
    gather8_kernel<<<148 * 16, 256>>>((const double*)table, idx, nidx, (double*)out);
    cudaEventRecord(a); gather8_kernel<<<148 * 16, 256>>>((const double*)table, idx, nidx, (double*)out); cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    printf("gather 8 B from 8 MB (L2): %.1f G elem/s (%.1f GB/s of 32 B sectors, idx stream %.1f GB/s)\n",
           nidx / time_ms(a, b) / 1e6, nidx * 32.0 / time_ms(a, b) / 1e6, nidx * 4.0 / time_ms(a, b) / 1e6);
  }
  return 0;
}
