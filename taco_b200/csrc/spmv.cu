// spmv.cu -- y(i) = A(i,j) * x(j), A CSR, fp32 / fp64.
//
// Replaces the CUDA the reference emits for SpMV (schedules scheduleSpMVGPU /
// scheduleSpMVSplitPosGPU / scheduleSpMVRowsGPU, /root/reference/test/tests-scheduling-eval.cpp:193-247):
//   reference: nnz-split over fused pos space, per-thread binary search, ONE GLOBAL fp64 atomicAdd PER NONZERO,
//              host-serial zeroing of y, block-start array allocated+freed per call in managed memory.
//   here     : ONE kernel.  The nonzeros are cut into tiles of TILE = 2048; CTA b
//              (0) immediately issues aligned 128-bit streaming loads of its window of crd / vals (they depend on
//                  nothing but blockIdx) plus a 64-entry overlap into the next window,
//              (1) meanwhile finds the rows whose first nonzero lies in the window with two 128-ary searches over pos
//                  (3 rounds for 1M rows; the search of taco_binarySearchBeforeBlock,
//                  /root/reference/src/codegen/codegen_cuda.cpp:110-125, without the block-start array and its
//                  extra launch),
//              (2) gathers x through L2 (evict_last), stages the products in shared memory,
//              (3) reduces every owned row with one thread in ascending position order -- the operation order of the
//                  reference's C kernel (Appendix A.1), so y is bit-identical to the oracle.  A row has exactly one
//                  owner: no atomics, no zero-fill pass.
//              Rows that run past the overlap are finished by the tile that contains their end: the owner publishes
//              the in-order sum of its piece (partial + epoch flag), tiles lying wholly inside the row publish a tree
//              sum of their window, and the end tile -- which only ever waits for LOWER-numbered tiles, so in-order CTA
//              dispatch guarantees progress -- adds the carried value and its own piece in order.  Rows spanning at
//              most two tiles (<= 2048 nonzeros always do) therefore stay bit-exact; longer rows are deterministic and
//              within 1e-12 of the sequential sum.
// Algorithmic bytes per launch: nnz*(4+sizeof T) + 4(n+1) + sizeof T*(cols + rows).
#include <climits>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace tb {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_VEC = 4;                 // nonzeros per thread per vector step (one int4 of crd)
constexpr int SPMV_STEPS = 2;
constexpr int SPMV_TILE = SPMV_THREADS * SPMV_VEC * SPMV_STEPS;     // 2048
constexpr int SPMV_OV = 64;                 // overlap into the next tile (rows ending inside it need no hand-over)

template <typename T> struct ValVec;
template <> struct ValVec<double> {
  static __device__ __forceinline__ void load4(const double* p, double (&v)[4]) {
    double2 a = tbd::ldg_stream_d2(p), b = tbd::ldg_stream_d2(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
};
template <> struct ValVec<float> {
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 a = tbd::ldg_stream_f4(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};

// `vec`: crd + q and vals + q are 16-byte aligned for every q this launch asks for (see `shift` in the kernel)
template <typename T>
__device__ __forceinline__ void spmv_load4(const int* __restrict__ crd, const T* __restrict__ vals, int q, int nnz, bool vec, int4& c,
                                           T (&v)[4]) {
  if (vec && q >= 0 && q + SPMV_VEC <= nnz) {
    c = tbd::ldg_stream_i4(crd + q);
    ValVec<T>::load4(vals + q, v);
  } else {
    int cc[4];
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const bool ok = q + e >= 0 && q + e < nnz;
      cc[e] = ok ? __ldg(crd + q + e) : 0;
      v[e] = ok ? __ldg(vals + q + e) : T(0);
    }
    c = make_int4(cc[0], cc[1], cc[2], cc[3]);
  }
}

// GV = 0: L1-allocating gather (default), GV = 1: L1::no_allocate (the gathered sectors of a uniform matrix are never re-used
// inside an SM; measured as TACO_B200_SPMV_VARIANT=3)
template <typename T, int GV = 0>
__device__ __forceinline__ T spmv_ld_x(const T* p, uint64_t keep) {
  T r;
  if constexpr (GV == 1) {
    if constexpr (sizeof(T) == 8) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(keep));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(keep));
  } else {
    if constexpr (sizeof(T) == 8) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(keep));
    else asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(keep));
  }
  return r;
}

// deterministic tree sum of prod[a..b) by the whole CTA; result valid in thread 0
template <typename T>
__device__ __forceinline__ T spmv_block_sum(const T* prod, int a, int b, T* red) {
  T acc = T(0);
  for (int q = a + (int)threadIdx.x; q < b; q += SPMV_THREADS) acc += prod[q];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  T tot = T(0);
  if (threadIdx.x == 0)
    for (int w = 0; w < SPMV_THREADS / 32; w++) tot += red[w];
  return tot;
}

// MAPPED: row r is stored at y[ymap[r]] instead of y[r] (TTV: the rows are the fibers of a CSF tensor and ymap holds their
// positions in the dense result, csf.cu); the unmapped instantiation is unchanged by the flag.
template <typename T, int MINB, bool MAPPED, int GV = 0, int STEPS = SPMV_STEPS>
__global__ void __launch_bounds__(SPMV_THREADS, MINB)
spmv_csr_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                const T* __restrict__ x, T* __restrict__ y, int rows, int nnz, T* __restrict__ partial,
                int* __restrict__ flag, int epoch, const unsigned* __restrict__ ymap, int shift, bool vec) {
  constexpr int TILE = SPMV_THREADS * SPMV_VEC * STEPS;      // nonzeros per CTA (TACO_B200_SPMV_VARIANT=4..6 sweep it)
  __shared__ T prod[TILE + SPMV_OV];
  __shared__ T red[SPMV_THREADS / 32];
  __shared__ int s_cnt[SPMV_THREADS / 32];
  __shared__ int s_long[4];                        // tail row, its start
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  // Tiles start `shift` (0..3) nonzeros before the arrays do, so that every 128-bit load of crd / vals is aligned even when the
  // caller hands over a slice that starts at an arbitrary element (a row shard of a larger matrix): tile 0 is short.
  const int lo = b * TILE - shift;            // window [lo, hi), staged [lo, hiov); positions < 0 do not exist
  const int hi = min(lo + TILE, nnz);
  const int hiov = min(lo + TILE + SPMV_OV, nnz);
  const bool last_tile = (b == (int)gridDim.x - 1);

  // ---- (0) window loads: independent of the row structure ------------------------------------------------------
  int4 c[STEPS];
  T v[STEPS][4];
#pragma unroll
  for (int s = 0; s < STEPS; s++) spmv_load4<T>(crd, vals, lo + (s * SPMV_THREADS + tid) * SPMV_VEC, nnz, vec, c[s], v[s]);
  if (tid == 0) s_long[0] = -1;

  // ---- (1) owned rows [r_lo, r_hi): first row with pos[r] >= lo / >= hi.  Warp 0 alone runs two 16-ary searches
  //          (half a warp each, 5 rounds for 1M rows, no CTA barrier) while the other warps go on to the gathers.
  if (tid < 32) {
    const int half = tid >> 4, t = tid & 15;
    const unsigned hmask = half ? 0xffff0000u : 0x0000ffffu;
    const int target = half ? hi : lo;
    int base = 0, n = rows + 1;                    // invariant: pos[base + n - 1] >= target  (pos[rows] = nnz)
    for (int nmax = rows + 1; nmax > 1; nmax = (nmax + 15) >> 4) {   // same trip count for both halves
      const int stride = (n + 15) >> 4;
      const int idx = base + t * stride;
      const bool below = idx < base + n && __ldg(pos + idx) < target;
      const int f = __popc(__ballot_sync(0xffffffffu, below) & hmask);
      if (f == 0) { n = 1; }
      else {
        const int nb = base + (f - 1) * stride + 1;
        n = min(stride, base + n - nb);
        base = nb;
      }
    }
    if (t == 0) s_cnt[half] = base;
    // pull the pos entries of the owned rows towards L2 while the gathers run (they are needed after the barrier)
    const int p_lo = __shfl_sync(0xffffffffu, base, 0), p_hi = last_tile ? rows : __shfl_sync(0xffffffffu, base, 16);
    if (p_lo + tid * 32 <= p_hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(pos + p_lo + tid * 32));
  }

  // ---- (2) products into shared memory ------------------------------------------------------------------------------
  const uint64_t keep = tbd::policy_evict_last();
#pragma unroll
  for (int s = 0; s < STEPS; s++) {
    const int q = (s * SPMV_THREADS + tid) * SPMV_VEC;
    if (lo + q < hiov) {
      const T x0 = spmv_ld_x<T, GV>(x + c[s].x, keep), x1 = spmv_ld_x<T, GV>(x + c[s].y, keep), x2 = spmv_ld_x<T, GV>(x + c[s].z, keep),
              x3 = spmv_ld_x<T, GV>(x + c[s].w, keep);
      T* d = prod + q;
      d[0] = v[s][0] * x0; d[1] = v[s][1] * x1; d[2] = v[s][2] * x2; d[3] = v[s][3] * x3;
    }
  }
  // the overlap into the next window: 16 threads of warp 1
  if (tid >= 32 && tid < 32 + SPMV_OV / SPMV_VEC && lo + TILE + (tid - 32) * SPMV_VEC < hiov) {
    const int q = TILE + (tid - 32) * SPMV_VEC;
    spmv_load4<T>(crd, vals, lo + q, nnz, vec, c[0], v[0]);
    const T x0 = spmv_ld_x<T, GV>(x + c[0].x, keep), x1 = spmv_ld_x<T, GV>(x + c[0].y, keep), x2 = spmv_ld_x<T, GV>(x + c[0].z, keep),
            x3 = spmv_ld_x<T, GV>(x + c[0].w, keep);
    T* d = prod + q;
    d[0] = v[0][0] * x0; d[1] = v[0][1] * x1; d[2] = v[0][2] * x2; d[3] = v[0][3] * x3;
  }
  __syncthreads();
  const int r_lo = s_cnt[0];
  const int r_hi = last_tile ? rows : s_cnt[1];

  // ---- (3) owned rows: one thread per row, ascending positions ------------------------------------------------------
  for (int r = r_lo + tid; r < r_hi; r += SPMV_THREADS) {
    const int s = __ldg(pos + r), e = __ldg(pos + r + 1);
    if (e <= hiov) {
      T acc = T(0);
      for (int q = s - lo; q < e - lo; q++) acc += prod[q];
      if constexpr (MAPPED) y[__ldg(ymap + r)] = acc;
      else y[r] = acc;
    } else {                                       // only the last owned row can run past the overlap
      T acc = T(0);
      for (int q = s - lo; q < hi - lo; q++) acc += prod[q];
      partial[b] = acc;                            // in-order sum of the owner's piece
      __threadfence();
      atomicExch(flag + b, epoch);
    }
  }

  // ---- (4) the row that covers nonzero `lo` but started in an earlier tile -----------------------------------------------
  // The tile in which such a row ENDS waits for the partial sums of the lower-numbered tiles it runs through.  This relies on
  // the CTAs of a 1-D grid being dispatched in blockIdx order (a lower tile is running or done whenever a higher one is
  // resident).  An arrival-order ticket would remove the assumption but measured 64 -> 70 us on config C1
  // (profiles/r02_variants.md); spmv_warp_kernel below has no cross-CTA wait at all (TACO_B200_SPMV_KERNEL=1).
  if (r_lo > 0 && lo < nnz) {
    const int e_head = __ldg(pos + r_lo);          // uniform over the CTA
    if (e_head > lo) {
      const int s_head = __ldg(pos + r_lo - 1);
      const int b0 = (s_head + shift) / TILE; // owner tile
      if (e_head > min((b0 + 1) * TILE - shift + SPMV_OV, nnz)) {     // the owner handed this row over
        if (e_head > hi) {                         // this window lies wholly inside the row
          const T sum = spmv_block_sum(prod, 0, hi - lo, red);
          if (tid == 0) {
            partial[b] = sum;
            __threadfence();
            atomicExch(flag + b, epoch);
          }
        } else if (tid == 0) {                     // the row ends here: carry the earlier pieces in tile order
          T acc = T(0);
          for (int bb = b0; bb < b; bb++) {
            while (atomicAdd(flag + bb, 0) != epoch) __nanosleep(40);
            __threadfence();
            const T pv = *(volatile T*)(partial + bb);
            acc = (bb == b0) ? pv : acc + pv;
          }
          for (int q = 0; q < e_head - lo; q++) acc += prod[q];
          if constexpr (MAPPED) y[__ldg(ymap + r_lo - 1)] = acc;
          else y[r_lo - 1] = acc;
        }
      }
    }
  }
}

// =========================================================================================================
// Barrier-free variant (TTV's default; SpMV keeps the CTA-tile kernel above, see spmv_launch_raw): one WARP owns a chunk of SW_CHUNK consecutive nonzeros and walks it in windows of
// SW_WIN = 128 (4 per lane, one 128-bit load of crd and two of vals per lane).  Per window: products into the warp's own
// 1 KB shared-memory slice (__syncwarp only -- no CTA barrier anywhere), then
//   * the row left open by the previous window is continued in order (its running sum lives in a register),
//   * every row that STARTS in the window is found through a row cursor (32 pos entries per step, ballot), and one lane
//     sums it in ascending position order -- the reference C kernel's order, bit-identical to it; the row that runs past
//     the window becomes the open row.
// A row still open at the end of the chunk is finished by the same warp (it reads on past its chunk; the next chunk's
// warp starts at the first row that starts in ITS chunk and skips what lies before it).  Rows longer than SW_LONG are not
// summed here at all: their start lane appends them to a list and a second, tiny kernel sums each with a fixed-shape
// strided + tree reduction (deterministic, within 1e-12 of the sequential order).  Windows that lie wholly inside such a
// row are skipped without loading anything.  No flags, no epochs, no dependence on CTA dispatch order; 64 independent
// warps per SM keep the L1 gather pipe busy where the CTA-wide phases of the tile kernel left it idle a third of the time.
// =========================================================================================================
constexpr int SW_WIN = 128;
constexpr int SW_LONG = 1024;
constexpr int SW_WARPS = 8;

// first i in [0, n) with a[i] >= target (a non-decreasing), n if none: 32-ary search by one warp, 4-5 dependent rounds
__device__ __forceinline__ int warp_first_ge(const int* __restrict__ a, int n, int target, int lane) {
  int base = 0, len = n;
  while (len > 0) {
    const int stride = (len + 31) >> 5;
    const int idx = base + lane * stride;
    const bool below = idx < base + len && __ldg(a + idx) < target;
    const int f = __popc(__ballot_sync(0xffffffffu, below));
    if (f == 0) return base;
    const int nb = base + (f - 1) * stride + 1;
    const int ne = min(base + f * stride, base + len);
    base = nb;
    len = ne - nb;
  }
  return base;
}

template <typename T, bool MAPPED, int MINB, int WINS>
__global__ void __launch_bounds__(SW_WARPS * 32, MINB)
spmv_warp_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals, const T* __restrict__ x,
                 T* __restrict__ y, int rows, int nnz, const unsigned* __restrict__ ymap, int shift, bool vec, int nchunks,
                 int* __restrict__ long_rows, int* __restrict__ counters, int long_cap) {
  __shared__ T prod_all[SW_WARPS][SW_WIN];
  const int lane = threadIdx.x & 31;
  T* prod = prod_all[threadIdx.x >> 5];
  const int w = blockIdx.x * SW_WARPS + (threadIdx.x >> 5);
  if (w >= nchunks) return;
  constexpr int SW_CHUNK = SW_WIN * WINS;          // nonzeros per warp
  const int lo = w * SW_CHUNK - shift;             // chunks start `shift` nonzeros early so that 128-bit loads stay aligned
  const int hi = min(lo + SW_CHUNK, nnz);
  const uint64_t keep = tbd::policy_evict_last();
  // row cursor: the first row that starts in this chunk
  int r = warp_first_ge(pos, rows + 1, max(lo, 0), lane);
  int next_start = r < rows ? __ldg(pos + r) : INT_MAX;
  bool open = false;
  int open_row = 0, open_end = 0;
  T acc = T(0);
  for (int p0 = lo;; p0 += SW_WIN) {
    const bool in_chunk = p0 < hi;
    if (!in_chunk && !open) break;
    const int p1 = min(p0 + SW_WIN, in_chunk ? hi : open_end);
    if (!open && next_start >= p1) continue;       // nothing starts or continues here (inside a long / foreign row)
    // ---- products of the window into the warp's slice ----------------------------------------------------------
    {
      const int q = p0 + lane * SPMV_VEC;
      if (q < p1) {
        int4 c;
        T v[4];
        spmv_load4<T>(crd, vals, q, nnz, vec, c, v);
        const T x0 = spmv_ld_x<T, 0>(x + c.x, keep), x1 = spmv_ld_x<T, 0>(x + c.y, keep), x2 = spmv_ld_x<T, 0>(x + c.z, keep),
                x3 = spmv_ld_x<T, 0>(x + c.w, keep);
        T* d = prod + lane * SPMV_VEC;
        d[0] = v[0] * x0; d[1] = v[1] * x1; d[2] = v[2] * x2; d[3] = v[3] * x3;
      }
    }
    __syncwarp();
    // ---- the row left open by an earlier window: continue in position order ---------------------------------------
    if (open) {
      const int e = min(open_end, p1);
      for (int i = max(p0, 0); i < e; i++) acc += prod[i - p0];
      if (open_end <= p1) {
        if (lane == 0) {
          if constexpr (MAPPED) y[__ldg(ymap + open_row)] = acc;
          else y[open_row] = acc;
        }
        open = false;
      }
    }
    // ---- rows that start in this window ---------------------------------------------------------------------------
    if (in_chunk) {
      while (next_start < p1) {
        const int ri = r + lane;
        int s = INT_MAX, e = INT_MAX;
        if (ri < rows) { s = __ldg(pos + ri); e = __ldg(pos + ri + 1); }
        const bool valid = ri < rows && s < p1;
        const int nv = __popc(__ballot_sync(0xffffffffu, valid));      // a prefix of the lanes
        const int deg = valid ? e - s : 0;
        if (valid) {
          if (deg > SW_LONG) {
            const int at = atomicAdd(counters, 1);
            if (at < long_cap) long_rows[at] = ri;
          } else if (e <= p1) {                      // complete inside the window (empty rows included)
            T a = T(0);
            for (int i = s; i < e; i++) a += prod[i - p0];
            if constexpr (MAPPED) y[__ldg(ymap + ri)] = a;
            else y[ri] = a;
          }
        }
        const unsigned em = __ballot_sync(0xffffffffu, valid && deg <= SW_LONG && e > p1);
        if (em) {                                    // at most one row runs past the window: it becomes the open row
          const int src = __ffs(em) - 1;
          T a = T(0);
          if (lane == src)
            for (int i = s; i < p1; i++) a += prod[i - p0];
          acc = __shfl_sync(0xffffffffu, a, src);
          open_row = __shfl_sync(0xffffffffu, ri, src);
          open_end = __shfl_sync(0xffffffffu, e, src);
          open = true;
        }
        const int s_next = __shfl_sync(0xffffffffu, nv == 32 ? e : s, nv == 32 ? 31 : nv);
        r += nv;
        next_start = r < rows ? s_next : INT_MAX;
        if (nv == 0) break;
      }
    }
    __syncwarp();                                    // the slice is rewritten by the next window
  }
  // trailing rows that start at nnz (all empty) belong to the last chunk
  if (w == nchunks - 1)
    for (int rr = r + lane; rr < rows; rr += 32) {
      if constexpr (MAPPED) y[__ldg(ymap + rr)] = T(0);
      else y[rr] = T(0);
    }
}

// rows longer than SW_LONG: one CTA per row, fixed strided partial sums + fixed tree -> deterministic.  The last CTA to
// finish resets the list counters for the next launch on this stream.
template <typename T, bool MAPPED>
__global__ void __launch_bounds__(256)
spmv_long_rows_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals, const T* __restrict__ x,
                      T* __restrict__ y, const unsigned* __restrict__ ymap, const int* __restrict__ long_rows, int* __restrict__ counters,
                      int long_cap) {
  __shared__ T red[256];
  const int tid = threadIdx.x;
  const int count = min(*(volatile int*)counters, long_cap);
  const uint64_t keep = tbd::policy_evict_last();
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    const int r = long_rows[i];
    const int s = __ldg(pos + r), e = __ldg(pos + r + 1);
    T a = T(0);
    for (int p = s + tid; p < e; p += 256) a += __ldg(vals + p) * spmv_ld_x<T, 0>(x + __ldg(crd + p), keep);
    red[tid] = a;
    __syncthreads();
#pragma unroll
    for (int off = 128; off > 0; off >>= 1) {
      if (tid < off) red[tid] += red[tid + off];
      __syncthreads();
    }
    if (tid == 0) {
      if constexpr (MAPPED) y[__ldg(ymap + r)] = red[0];
      else y[r] = red[0];
    }
    __syncthreads();
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(counters + 1, 1) == (int)gridDim.x - 1) { counters[0] = 0; counters[1] = 0; __threadfence(); }
  }
}

// Hand-over scratch (one partial + one flag per tile).  Flags are compared against a per-launch epoch, so they are never
// cleared between calls.  One scratch set per STREAM: launches on one stream are serialised by the stream, launches on
// different streams (taco_b200_set_stream) or from different host threads never share slots; the table is mutex-guarded.
struct SpmvScratch {
  void* partial = nullptr; int* flag = nullptr; int cap = 0; int epoch = 0;      // tile kernel: hand-over slots
  int* long_rows = nullptr; int* counters = nullptr; int long_cap = 0;            // warp kernel: long-row list + {count, done}
};
static std::mutex g_spmv_mu;
static std::unordered_map<cudaStream_t, SpmvScratch> g_spmv_scratch;

static int spmv_long_scratch_for(cudaStream_t st, int cap, SpmvScratch* out) {
  std::lock_guard<std::mutex> lk(g_spmv_mu);
  SpmvScratch& sc = g_spmv_scratch[st];
  if (!sc.counters) {
    TB_CUDA(cudaMallocAsync((void**)&sc.counters, sizeof(int) * 4, st));
    TB_CUDA(cudaMemsetAsync(sc.counters, 0, sizeof(int) * 4, st));     // the long-row kernel leaves them zeroed again
  }
  if (cap > sc.long_cap) {
    if (sc.long_rows) cudaFreeAsync(sc.long_rows, st);
    sc.long_cap = cap + cap / 2 + 1024;
    sc.long_rows = nullptr;
    TB_CUDA(cudaMallocAsync((void**)&sc.long_rows, sizeof(int) * (size_t)sc.long_cap, st));
  }
  *out = sc;
  return TACO_B200_OK;
}

static int spmv_scratch_for(cudaStream_t st, int ntiles, SpmvScratch* out) {
  std::lock_guard<std::mutex> lk(g_spmv_mu);
  SpmvScratch& sc = g_spmv_scratch[st];
  if (ntiles > sc.cap) {
    if (sc.partial) { cudaFreeAsync(sc.partial, st); cudaFreeAsync(sc.flag, st); }     // stream-ordered: earlier launches finish first
    sc.cap = ntiles + ntiles / 2 + 1024;
    sc.partial = nullptr; sc.flag = nullptr;
    TB_CUDA(cudaMallocAsync(&sc.partial, sizeof(double) * (size_t)sc.cap, st));
    TB_CUDA(cudaMallocAsync((void**)&sc.flag, sizeof(int) * (size_t)sc.cap, st));
    TB_CUDA(cudaMemsetAsync(sc.flag, 0, sizeof(int) * (size_t)sc.cap, st));
    sc.epoch = 0;
  }
  if (++sc.epoch == INT32_MAX) {
    TB_CUDA(cudaMemsetAsync(sc.flag, 0, sizeof(int) * (size_t)sc.cap, st));
    sc.epoch = 1;
  }
  *out = sc;
  return TACO_B200_OK;
}

template <typename T>
static int spmv_launch_raw(const int* pos, const int* crd, const T* vals, const T* x, T* y, int rows, int nnz, const unsigned* ymap,
                           const char* prof_name) {
  static const int variant = getenv("TACO_B200_SPMV_VARIANT") ? atoi(getenv("TACO_B200_SPMV_VARIANT")) : 0;
  const int steps = (!ymap && (variant == 4 || variant == 6)) ? 1 : (!ymap && variant == 5) ? 4 : SPMV_STEPS;
  const int tile = SPMV_THREADS * SPMV_VEC * steps;
  // 128-bit loads need crd + q and vals + q 16-byte aligned at every vector start q = 4m - shift: possible when both arrays
  // are element-aligned and misaligned by the same number of elements modulo 4 (always true for a slice [o, o + nnz) of
  // aligned arrays, i.e. a row shard); anything else takes scalar loads
  const uintptr_t ca = (uintptr_t)crd, va = (uintptr_t)vals;
  int shift = 0;
  bool vec = (ca % 4 == 0) && (va % sizeof(T) == 0);
  if (vec) {
    shift = (int)((ca / 4) % 4);
    vec = ((va / sizeof(T)) - (uintptr_t)shift) % (16 / sizeof(T)) == 0;
    if (!vec) shift = 0;
  }
  // Which kernel: measured on the B200 (profiles/r02_variants.md) -- the CTA-tile kernel wins on CSR SpMV with rows of ~10
  // nonzeros (C1: 64.6 us against 80-89 us), the barrier-free warp-chunk kernel on TTV, whose "rows" (fibers) average 1.5 leaves
  // (0.710 ms against 0.738 ms).  TACO_B200_SPMV_KERNEL = 0 / 1..4 forces one or the other.
  static const int forced = getenv("TACO_B200_SPMV_KERNEL") ? atoi(getenv("TACO_B200_SPMV_KERNEL")) : -1;
  const int kernel_sel = forced >= 0 ? forced : (ymap ? 1 : 0);
  if (kernel_sel != 0) {
    // nonzeros per warp: 8 windows of 128 by default (TACO_B200_SPMV_KERNEL = 2: 16 windows, 3: 8 windows / 6 CTAs per SM,
    // 4: 4 windows)
    const int wins = kernel_sel == 2 ? 16 : kernel_sel == 4 ? 4 : 8;
    const int chunk = SW_WIN * wins;
    const int nchunks = nnz > 0 ? (int)(((long long)nnz + shift + chunk - 1) / chunk) : 1;
    const int lcap = nnz / (SW_LONG + 1) + 1;
    SpmvScratch sc;
    TB_TRY(spmv_long_scratch_for(stream(), lcap, &sc));
    const int grid = (nchunks + SW_WARPS - 1) / SW_WARPS;
    {
      ProfScope ps(prof_name);
#define TB_SW_GO(MAPPED, MINB, WINS)                                                                                                    \
  spmv_warp_kernel<T, MAPPED, MINB, WINS><<<grid, SW_WARPS * 32, 0, stream()>>>(pos, crd, vals, x, y, rows, nnz, ymap, shift, vec, nchunks, \
                                                                                sc.long_rows, sc.counters, lcap)
      if (ymap) {
        if (kernel_sel == 2) TB_SW_GO(true, 8, 16); else if (kernel_sel == 3) TB_SW_GO(true, 6, 8); else if (kernel_sel == 4) TB_SW_GO(true, 8, 4); else TB_SW_GO(true, 8, 8);
      } else {
        if (kernel_sel == 2) TB_SW_GO(false, 8, 16); else if (kernel_sel == 3) TB_SW_GO(false, 6, 8); else if (kernel_sel == 4) TB_SW_GO(false, 8, 4); else TB_SW_GO(false, 8, 8);
      }
#undef TB_SW_GO
      const int lgrid = lcap < num_sms() * 2 ? lcap : num_sms() * 2;
      if (ymap) spmv_long_rows_kernel<T, true><<<lgrid, 256, 0, stream()>>>(pos, crd, vals, x, y, ymap, sc.long_rows, sc.counters, lcap);
      else spmv_long_rows_kernel<T, false><<<lgrid, 256, 0, stream()>>>(pos, crd, vals, x, y, ymap, sc.long_rows, sc.counters, lcap);
    }
    count_launch(2);
    TB_CUDA(cudaGetLastError());
    return TACO_B200_OK;
  }
  const int ntiles = nnz > 0 ? (int)(((long long)nnz + shift + tile - 1) / tile) : 1;
  SpmvScratch sc;
  TB_TRY(spmv_scratch_for(stream(), ntiles, &sc));
  {
    ProfScope ps(prof_name);
#define TB_SPMV_GO(MINB, MAPPED, GV, ...)                                                                              \
  spmv_csr_kernel<T, MINB, MAPPED, GV, ##__VA_ARGS__><<<ntiles, SPMV_THREADS, 0, stream()>>>(pos, crd, vals, x, y, rows, nnz, \
                                                                              (T*)sc.partial, sc.flag, sc.epoch, ymap, shift, vec)
    if (ymap && variant == 3) TB_SPMV_GO(6, true, 1);
    else if (ymap) TB_SPMV_GO(6, true, 0);
    else if (variant == 1) TB_SPMV_GO(8, false, 0);
    else if (variant == 2) TB_SPMV_GO(7, false, 0);
    else if (variant == 3) TB_SPMV_GO(6, false, 1);
    else if (variant == 4) TB_SPMV_GO(6, false, 0, 1);       // 1024-nonzero tiles
    else if (variant == 5) TB_SPMV_GO(6, false, 0, 4);       // 4096-nonzero tiles
    else if (variant == 6) TB_SPMV_GO(8, false, 0, 1);       // 1024-nonzero tiles, 8 CTAs per SM
    else TB_SPMV_GO(6, false, 0);
#undef TB_SPMV_GO
  }
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

template <typename T>
static int spmv_launch(const CsrView& A, const In& pos, const In& crd, const In& vals, const In& x, Out& y, int nnz) {
  return spmv_launch_raw<T>(pos.as<int>(), crd.as<int>(), vals.as<T>(), x.as<T>(), y.as<T>(), A.rows, nnz, nullptr, "spmv_csr");
}

// csf.cu: y[ymap[r]] = sum_p vals[p] * x[crd[p]] over the "rows" r of any (pos, crd, vals) level (TTV over the fibers)
int spmv_mapped(DType dt, const int* pos, const int* crd, const void* vals, const void* x, void* y, int rows, int nnz,
                const unsigned* ymap, const char* prof_name) {
  if (dt == DType::F64) return spmv_launch_raw<double>(pos, crd, (const double*)vals, (const double*)x, (double*)y, rows, nnz, ymap, prof_name);
  return spmv_launch_raw<float>(pos, crd, (const float*)vals, (const float*)x, (float*)y, rows, nnz, ymap, prof_name);
}

int csr_nnz(const CsrView& A, int32_t vals_size_hint, int32_t* nnz) {
  if (!A.pos) return fail(TACO_B200_ERR_ARG, "CSR operand has no pos array");
  if (vals_size_hint > 0 && trusts_vals_size(A.pos)) { *nnz = vals_size_hint; return TACO_B200_OK; }
  return read_i32(A.pos + A.rows, nnz);
}

static int spmv_views(taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x, DenseView* yv, CsrView* Av, DenseView* xv) {
  TB_TRY(ensure_init());
  TB_TRY(view_dense(y, 1, "y", yv));
  TB_TRY(view_csr(A, "A", Av));
  TB_TRY(view_dense(x, 1, "x", xv));
  if (yv->dim[0] != Av->rows || xv->dim[0] != Av->cols)
    return fail(TACO_B200_ERR_ARG, "spmv: dimension mismatch y[%d] = A[%d x %d] * x[%d]", yv->dim[0], Av->rows, Av->cols,
                xv->dim[0]);
  if (yv->dt != Av->dt || xv->dt != Av->dt) return fail(TACO_B200_ERR_FORMAT, "spmv: mixed component types");
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

// assemble: dense result -> allocate vals (the reference's generated assemble does `y_vals = malloc(...)`).
int taco_b200_spmv_assemble(taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x) {
  DenseView yv, xv; CsrView Av;
  TB_TRY(spmv_views(y, A, x, &yv, &Av, &xv));
  void* p = result_alloc(yv.count() * dsize(yv.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "spmv: cannot allocate result");
  y->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

int taco_b200_spmv_compute(taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x) {
  DenseView yv, xv; CsrView Av;
  TB_TRY(spmv_views(y, A, x, &yv, &Av, &xv));
  int32_t nnz = 0;
  TB_TRY(csr_nnz(Av, A->vals_size, &nnz));
  if (nnz < 0) return fail(TACO_B200_ERR_ARG, "spmv: negative nnz");
  if (nnz > INT32_MAX - 65536) return fail(TACO_B200_ERR_ARG, "spmv: nnz too close to the int32 limit");
  size_t es = dsize(Av.dt);
  In pos, crd, vals, xin; Out yout;
  TB_TRY(pos.acquire(Av.pos, sizeof(int32_t) * ((size_t)Av.rows + 1)));
  TB_TRY(crd.acquire(Av.crd ? (void*)Av.crd : (void*)Av.pos, sizeof(int32_t) * (size_t)nnz));
  TB_TRY(vals.acquire(Av.vals ? Av.vals : (void*)Av.pos, es * (size_t)nnz));
  TB_TRY(xin.acquire(xv.vals, es * (size_t)Av.cols));
  TB_TRY(yout.acquire(yv.vals, es * (size_t)Av.rows));
  if (Av.rows > 0) {
    if (Av.dt == DType::F64) TB_TRY(spmv_launch<double>(Av, pos, crd, vals, xin, yout, nnz));
    else TB_TRY(spmv_launch<float>(Av, pos, crd, vals, xin, yout, nnz));
  }
  TB_TRY(yout.commit());
  return finish_call();
}

int taco_b200_spmv_evaluate(taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x) {
  TB_TRY(taco_b200_spmv_assemble(y, A, x));
  return taco_b200_spmv_compute(y, A, x);
}

}  // extern "C"
