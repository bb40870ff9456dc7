// scan.cuh -- device-wide exclusive prefix sum over int32, used to build `pos` from per-row counts.
//
// This is the GPU form of the reference's sequential pos construction in two-phase assembly
// (`C2_pos[0] = 0; for i: C2_pos[i+1] = C2_pos[i] + C2_nnz[i]`, emitted by lowerAssemble,
// /root/reference/src/lower/lowerer_impl_imperative.cpp:2616-2779 via CompressedModeFormat::getSeqInitEdges /
// getSeqInsertEdge, src/lower/mode_format_compressed.cpp:217-232) and of the append-mode finalize prefix sum
// (getAppendFinalizeLevel, mode_format_compressed.cpp:192-211).
// Reduce-then-scan, 3 launches per level: block-local scan of 2048 elements (warp shuffles), recursive scan of the
// block totals, uniform add.  in/out may alias.
#pragma once
#include "common.cuh"

namespace tb {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// out[i] = sum(in[0..i-1]) within the tile; totals[tile] = tile sum
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_kernel(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ totals, long long n) {
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS];
  int sum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += t;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    int ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int off = 1; off < SCAN_THREADS / 32; off <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, ws, off);
      if (lane >= off) ws += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = ws;
  }
  __syncthreads();
  int run = incl - sum + (wid > 0 ? warp_sums[wid - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == SCAN_THREADS - 1 && totals) totals[blockIdx.x] = run;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int* __restrict__ out, const int* __restrict__ offsets, long long n) {
  const int off = offsets[blockIdx.x];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) out[base + k] += off;
}

// 64-bit total of n int32 counts (one atomic per CTA): the size of a sparse result BEFORE its int32 prefix sum, so that a
// result with more than 2^31-1 entries is refused instead of wrapping negative and under-allocating crd / vals.
static __global__ void __launch_bounds__(256)
total_i64_kernel(const int* __restrict__ in, long long n, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long warp_sums[8];
  unsigned long long s = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += (unsigned)in[i];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < 8; w++) t += warp_sums[w];
    if (t) atomicAdd(total, t);
  }
}

// reads the total back (one synchronisation -- the caller needs the size on the host anyway) and refuses sizes beyond int32
static int checked_total_i32(const int* counts, long long n, const char* what, int32_t* total_out) {
  std::lock_guard<std::mutex> lk(small_scratch_mutex());
  void* d = small_scratch();
  if (!d) return fail(TACO_B200_ERR_ALLOC, "no device scratch");
  cudaError_t e = cudaMemsetAsync(d, 0, sizeof(unsigned long long), stream());
  if (e != cudaSuccess) return fail(TACO_B200_ERR_CUDA, "memset failed: %s", cudaGetErrorString(e));
  if (n > 0) {
    const long long ctas = (n + 2047) / 2048;
    total_i64_kernel<<<(unsigned)(ctas < 1184 ? ctas : 1184), 256, 0, stream()>>>(counts, n, (unsigned long long*)d);
    count_launch(1);
  }
  unsigned long long h = 0;
  TB_TRY(read_back(&h, d, sizeof(h)));
  if (h > (unsigned long long)INT32_MAX)
    return fail(TACO_B200_ERR_ARG, "%s: the result has %llu stored entries, more than int32 positions can address", what, h);
  *total_out = (int32_t)h;
  return TACO_B200_OK;
}

// exclusive scan of n int32 on the current stream; in and out may be the same buffer
static int exclusive_scan_i32(const int* in, int* out, long long n) {
  if (n <= 0) return TACO_B200_OK;
  const long long tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (tiles == 1) {
    scan_tile_kernel<<<1, SCAN_THREADS, 0, stream()>>>(in, out, nullptr, n);
    count_launch(1);
    TB_CUDA(cudaGetLastError());
    return TACO_B200_OK;
  }
  void* totals = nullptr;
  TB_TRY(scratch_alloc(&totals, sizeof(int) * (size_t)tiles));
  scan_tile_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, stream()>>>(in, out, (int*)totals, n);
  count_launch(1);
  int rc = exclusive_scan_i32((const int*)totals, (int*)totals, tiles);
  if (rc == TACO_B200_OK) {
    scan_add_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, stream()>>>(out, (const int*)totals, n);
    count_launch(1);
  }
  scratch_free(totals);
  TB_TRY(rc);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

}  // namespace tb
