// spmm.cu -- C(i,k) = A(i,j) * B(j,k), A CSR, B dense row-major, C dense (row-major or {1,0}), fp32 / fp64.
//
// Replaces the CUDA the reference emits for scheduleSpMMGPU (/root/reference/test/tests-scheduling-eval.cpp:249-268,
// SURVEY.md Appendix A.2):
//   reference: grid ceil(nnz/64), warp gets 8 nnz, lane <-> k%32, the binary search is redone for each of 4
//              `dense_val` passes, ONE GLOBAL atomicAdd PER (nnz, k) (8.2 G atomics at config C2), host-serial
//              zeroing of the 2 GB result in managed memory, K <= 128 only.
//   here     : two cooperating schedules, no atomics anywhere, every C row written exactly once.
//   (1) rows of at most LONG nonzeros -- nnz-balanced ROW-ALIGNED slots: warp w owns the rows whose first nonzero lies
//       in [w*W,(w+1)*W) (one binary search per slot in a pre-pass, the search of taco_binarySearchBeforeBlock,
//       /root/reference/src/codegen/codegen_cuda.cpp:110-125).  A lane owns 16 bytes of the dense row (4 fp32 / 2 fp64
//       columns), so every gathered row of B is one fully coalesced 512-byte warp load, accumulators live in registers
//       and each C row is stored once with a streaming 128-bit store.  Empty rows are zeroed by their owner.  Inside a
//       row the products are accumulated in ascending position order with separate multiply and add -- the reference
//       C kernel's order (Appendix A.1) -- so these rows are bit-identical to it.
//   (2) long rows (> LONG = 128 nonzeros; 70 % of the nonzeros of the power-law config C2) -- balanced ITEMS in COLUMN-PANEL
//       order.  The columns are cut into P panels; a long row's nonzeros inside one panel form work items of at most CAP
//       nonzeros (fused multiply-add: these rows are reassociated by the partial sums anyway).  Items are laid out
//       panel-major and handed to persistent warps in that order by a ticket, so at any moment the whole chip gathers rows
//       of B from one column window.  Each item stores its partial row sum; a combine kernel adds the partials of a row in
//       ascending column (= position) order.  Measured at C2 (profiles/r02_variants.md): what pays is the balanced item
//       decomposition with the lean staged inner loop (P = 1: 2.95 ms against 3.42 ms for row-owner warps with atomics on hub
//       rows); the panel order adds 10 % at P = 4 (2.68 ms) and LOSES beyond P = 8 -- a (row, panel) piece of a 129..512-nonzero
//       row becomes a handful of nonzeros whose 512-byte partial row costs more than the locality returns (P = 32: 3.63 ms).
//       The plan (long-row list, panel cuts by binary search, item slots by one warp-aggregated atomic per panel) is rebuilt
//       on the device every call: no cached inspector state, no host read-back, results independent of scheduling
//       (run-to-run deterministic, within 1e-5 / 1e-12 of the sequential order).
// Algorithmic bytes per launch: nnz*(4+sizeof T) + 4(n+1) + sizeof T*K*(cols + rows).
#include <climits>
#include <cstdlib>

#include "common.cuh"
#include "scan.cuh"

namespace tb {

constexpr int SPMM_W = 64;          // nonzeros per slot (one warp)

template <typename T, int VEC> struct Frag { T v[VEC]; };

// A lane's 16-byte piece of a gathered B row (L1-allocating: hot columns of a power-law matrix are re-used inside an SM).
template <typename T, int VEC>
__device__ __forceinline__ Frag<T, VEC> load_row(const T* __restrict__ p) {
  Frag<T, VEC> f;
  if constexpr (VEC == 4 && sizeof(T) == 4) {
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(f.v[0]), "=f"(f.v[1]), "=f"(f.v[2]), "=f"(f.v[3]) : "l"(p));
  } else if constexpr (VEC == 2 && sizeof(T) == 8) {
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(f.v[0]), "=d"(f.v[1]) : "l"(p));
  } else if constexpr (VEC == 2 && sizeof(T) == 4) {
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(f.v[0]), "=f"(f.v[1]) : "l"(p));
  } else {
    static_assert(VEC == 1, "unhandled fragment width");
    f.v[0] = __ldg(p);
  }
  return f;
}

// one component through the NVLink multicast mapping of the result (delivered to every GPU of the group)
template <typename T>
__device__ __forceinline__ void st_multicast(T* p, T v) {
  if constexpr (sizeof(T) == 4) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
  else asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <typename T, int VEC, bool COLMAJOR>
__device__ __forceinline__ void store_row_plain(T* __restrict__ C, size_t row, int col, int rows, int K, const Frag<T, VEC>& f) {
  if constexpr (COLMAJOR) {
#pragma unroll
    for (int e = 0; e < VEC; e++) C[(size_t)(col + e) * rows + row] = f.v[e];
  } else {
    T* p = C + row * K + col;
    if constexpr (VEC == 4 && sizeof(T) == 4) {
      asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(f.v[0]), "f"(f.v[1]), "f"(f.v[2]),
                   "f"(f.v[3]) : "memory");
    } else if constexpr (VEC == 2 && sizeof(T) == 8) {
      asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(f.v[0]), "d"(f.v[1]) : "memory");
    } else if constexpr (VEC == 2 && sizeof(T) == 4) {
      asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(f.v[0]), "f"(f.v[1]) : "memory");
    } else {
      static_assert(VEC == 1, "unhandled fragment width");
      p[0] = f.v[0];
    }
  }
}

// The row with its fan-out (common.cuh): fo.n == 0 local only; -1 one store through the multicast mapping; n > 0 the local
// store and one store into each of n peer GPUs' copies of the result.
template <typename T, int VEC, bool COLMAJOR>
__device__ __forceinline__ void store_row(T* __restrict__ C, size_t row, int col, int rows, int K, const Frag<T, VEC>& f,
                                          const Fanout& fo) {
  if (fo.n < 0) {
    const long long mcd = fo.d[0];
    if constexpr (COLMAJOR) {
#pragma unroll
      for (int e = 0; e < VEC; e++) st_multicast<T>((T*)((char*)(C + (size_t)(col + e) * rows + row) + mcd), f.v[e]);
    } else {
      T* p = (T*)((char*)(C + row * K + col) + mcd);
      if constexpr (VEC == 4 && sizeof(T) == 4) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(f.v[0]), "f"(f.v[1]), "f"(f.v[2]),
                     "f"(f.v[3]) : "memory");
      } else {
#pragma unroll
        for (int e = 0; e < VEC; e++) st_multicast<T>(p + e, f.v[e]);
      }
    }
    return;
  }
  store_row_plain<T, VEC, COLMAJOR>(C, row, col, rows, K, f);
  for (int i = 0; i < fo.n; i++) store_row_plain<T, VEC, COLMAJOR>((T*)((char*)C + fo.d[i]), row, col, rows, K, f);
}

// A launch covers the row range [r0, r1) = nonzeros [p0, p1) (the whole matrix, or one row chunk of the host-operand
// pipeline below).
struct SpmmRange { int r0, r1, p0, p1; };

// Pre-pass: slot_rows[w] = first row whose first nonzero is at or after p0 + w*W; slot_rows[nslots] = r1.
__global__ void spmm_slot_rows_kernel(const int* __restrict__ pos, SpmmRange rg, int nslots, int* __restrict__ slot_rows) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nslots) return;
  slot_rows[w] = (w == nslots) ? rg.r1 : tbd::search_first_ge(pos, rg.r0, rg.r1, rg.p0 + w * SPMM_W);
}

// One staged nonzero: column and value side by side, so that a warp reads both with ONE broadcast shared-memory load.
template <typename T> struct __align__(8) SpmmNz { int c; T v; };

// acc += sum over nonzeros p in [a,b) of vals[p] * B[crd[p], col..col+VEC), in ascending p.  32 nonzeros are fetched with one
// coalesced load per array and parked in the warp's shared-memory stage; per nonzero the warp then issues one broadcast LDS
// (column + value), one 64-bit multiply-add for the row address (unsigned column x row stride in bytes) and one 16-byte
// gather, U gathers in flight.  FMA = false keeps multiply and add separate -- the reference C kernel's arithmetic
// (Appendix A.1), bit-identical to it; FMA = true (long rows, whose partial sums are reassociated anyway) contracts them.
template <typename T, int VEC, int U, bool FMA>
__device__ __forceinline__ void spmm_accumulate(Frag<T, VEC>& acc, const int* __restrict__ crd, const T* __restrict__ vals,
                                                const char* __restrict__ Bbytes, unsigned stride, int a, int b, int lane,
                                                SpmmNz<T>* __restrict__ stage) {
  for (int pb = a; pb < b; pb += 32) {
    const int cnt = min(32, b - pb);
    __syncwarp();                       // the previous chunk has been consumed
    if (lane < cnt) {
      SpmmNz<T> e;
      e.c = tbd::ldg_stream_i32(crd + pb + lane);
      e.v = __ldg(vals + pb + lane);
      stage[lane] = e;
    }
    __syncwarp();
    int j = 0;
    for (; j + U <= cnt; j += U) {
      Frag<T, VEC> bv[U];
      T v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const SpmmNz<T> e = stage[j + u];
        v[u] = e.v;
        bv[u] = load_row<T, VEC>((const T*)(Bbytes + (unsigned long long)(unsigned)e.c * stride));
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
#pragma unroll
        for (int x = 0; x < VEC; x++) {
          if constexpr (FMA) acc.v[x] = sizeof(T) == 4 ? (T)__fmaf_rn((float)v[u], (float)bv[u].v[x], (float)acc.v[x]) : (T)__fma_rn((double)v[u], (double)bv[u].v[x], (double)acc.v[x]);
          else acc.v[x] = acc.v[x] + v[u] * bv[u].v[x];                       // mul then add: never fused (-fmad=false)
        }
      }
    }
    for (; j < cnt; j++) {              // 1 .. U-1 left
      const SpmmNz<T> e = stage[j];
      const Frag<T, VEC> b1 = load_row<T, VEC>((const T*)(Bbytes + (unsigned long long)(unsigned)e.c * stride));
#pragma unroll
      for (int x = 0; x < VEC; x++) {
        if constexpr (FMA) acc.v[x] = sizeof(T) == 4 ? (T)__fmaf_rn((float)e.v, (float)b1.v[x], (float)acc.v[x]) : (T)__fma_rn((double)e.v, (double)b1.v[x], (double)acc.v[x]);
        else acc.v[x] = acc.v[x] + e.v * b1.v[x];
      }
    }
  }
}

// The same sum with the (column, value) pairs broadcast by shuffle instead of staged in shared memory: one dependent step
// fewer between the crd load and the first gather.  The short-row kernel is latency-bound (rows average 10 nonzeros: one
// chunk, one chain crd -> gather -> store per row), so it takes this form (measured: 1.24 ms against 1.40 ms staged); the
// long-row kernel, issue- and L1-bound over runs of up to 256 nonzeros, takes the staged form (1.29 ms against 1.58 ms).
// The row address is the plain element form Bcol + (size_t)c * K here: the byte form with IMAD.WIDE.U32 that pays in the
// issue-bound long kernel measured SLOWER in this latency-bound one (same-box A/B: TTM 3.16 -> 2.47 ms, C2 2.85 -> 2.74 ms).
// Always separate multiply and add: the reference's arithmetic.
template <typename T, int VEC, int U>
__device__ __forceinline__ void spmm_accumulate_shfl(Frag<T, VEC>& acc, const int* __restrict__ crd, const T* __restrict__ vals,
                                                     const T* __restrict__ Bcol, int K, int a, int b, int lane) {
  for (int pb = a; pb < b; pb += 32) {
    const int cnt = min(32, b - pb);
    int my_c = 0;
    T my_v = T(0);
    if (lane < cnt) {
      my_c = tbd::ldg_stream_i32(crd + pb + lane);
      my_v = __ldg(vals + pb + lane);
    }
    int j = 0;
    for (; j + U <= cnt; j += U) {
      Frag<T, VEC> bv[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int c = __shfl_sync(0xffffffffu, my_c, j + u);
        bv[u] = load_row<T, VEC>(Bcol + (size_t)c * K);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const T v = __shfl_sync(0xffffffffu, my_v, j + u);
#pragma unroll
        for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + v * bv[u].v[x];   // mul then add: never fused
      }
    }
    if (j < cnt) {                      // 1 .. U-1 left: same two phases, warp-uniform predicates (no loop: at most U-1 gathers)
      const int rem = cnt - j;
      Frag<T, VEC> bv[U > 1 ? U - 1 : 1];
#pragma unroll
      for (int u = 0; u < U - 1; u++) {
        if (u < rem) {
          const int c = __shfl_sync(0xffffffffu, my_c, j + u);
          bv[u] = load_row<T, VEC>(Bcol + (size_t)c * K);
        }
      }
#pragma unroll
      for (int u = 0; u < U - 1; u++) {
        if (u < rem) {
          const T v = __shfl_sync(0xffffffffu, my_v, j + u);
#pragma unroll
          for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + v * bv[u].v[x];
        }
      }
    }
  }
}

// Schedule (1): the rows of at most `long_thresh` nonzeros.  RMAP: result row r is stored at row rowmap[r] of C (TTM: the
// rows are the fibers of a CSF tensor, rowmap their cells in the dense (i,j) plane, csf.cu).
// MC: the result rows fan out to the other GPUs (`fo_arg`, common.cuh); a template flag so that the common
// single-GPU instantiation does not carry the offset in registers (32-register budget, every spill is in the per-row path)
template <typename T, int VEC, bool COLMAJOR, int U, int WARPS, int MINB, bool RMAP = false, bool MC = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_csr_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                const T* __restrict__ B, T* __restrict__ C, int rows, int K, SpmmRange rg, int nslots,
                const int* __restrict__ slot_rows, int long_thresh, Fanout fo_arg, const unsigned* __restrict__ rowmap = nullptr) {
  Fanout fo;
  if constexpr (MC) fo = fo_arg;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (w >= nslots) return;
  const int lo = rg.p0 + w * SPMM_W;
  // the slot's own window of crd / vals is needed three dependent loads from now (slot_rows -> pos -> crd): pull it into L2
  // meanwhile.  Issued before anything is known about the slot -- slots inside a long row fetch 512 bytes for nothing (+0.4 GB
  // at C2), but placing it after the slot_rows load cost the latency-bound short-row kernels 15 % (TTM 2.50 -> 2.93 ms)
  if (lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(crd + lo + lane * 32));
  else if (lane < 2 + (int)(2 * sizeof(T) / 4)) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + lo + (lane - 2) * (128 / (int)sizeof(T))));
  const int R0 = __ldg(slot_rows + w), R1 = __ldg(slot_rows + w + 1);
  if (R1 <= R0) return;                 // the slot lies inside one row that started earlier (typically a long row)
  const int col = (blockIdx.y * 32 + lane) * VEC;
  const bool active = col < K;
  const T* Bcol = B + (active ? col : 0);        // inactive lanes (ragged K) gather column 0 and never store
  for (int rb = R0; rb < R1; rb += 32) {
    const int r = rb + lane;
    const bool valid = r < R1;
    const int s = valid ? __ldg(pos + r) : 0;
    const int e = valid ? __ldg(pos + r + 1) : 0;
    unsigned empty = __ballot_sync(0xffffffffu, valid && e == s);
    unsigned full = __ballot_sync(0xffffffffu, valid && e > s && e - s <= long_thresh);
    Frag<T, VEC> acc;
#pragma unroll
    for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
    while (empty) {
      const int h = __ffs(empty) - 1;
      empty &= empty - 1;
      if (active) store_row<T, VEC, COLMAJOR>(C, RMAP ? (size_t)__ldg(rowmap + rb + h) : (size_t)(rb + h), col, rows, K, acc, fo);
    }
    while (full) {
      const int h = __ffs(full) - 1;
      full &= full - 1;
      const int hs = __shfl_sync(0xffffffffu, s, h);
      const int he = __shfl_sync(0xffffffffu, e, h);
#pragma unroll
      for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
      spmm_accumulate_shfl<T, VEC, U>(acc, crd, vals, Bcol, K, hs, he, lane);
      if (active) store_row<T, VEC, COLMAJOR>(C, RMAP ? (size_t)__ldg(rowmap + rb + h) : (size_t)(rb + h), col, rows, K, acc, fo);
    }
  }
}


static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

// ---------------------------------------------------------------------------------------------------------
// Schedule (2): long rows, column-panel order.
// ---------------------------------------------------------------------------------------------------------
struct SpmmLongCfg {
  int thresh;       // rows with more nonzeros than this are long
  int panels;       // column panels (<= 32)
  int panel_w;      // columns per panel
  int cap;          // nonzeros per work item
  int nlong_max;    // capacity of the long-row list: (p1 - p0) / (thresh + 1)
  long long items_max;
};

// counters[0] = number of long rows; counters[1] = ticket of the item kernel (both zeroed by a memset node)
constexpr int SPMM_FIND_ROWS = 4;      // rows per thread
__global__ void __launch_bounds__(256)
spmm_long_find_kernel(const int* __restrict__ pos, SpmmRange rg, SpmmLongCfg cfg, int* __restrict__ long_rows, int* __restrict__ counters) {
  __shared__ int s_cnt, s_base;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const long long base = (long long)rg.r0 + ((long long)blockIdx.x * blockDim.x) * SPMM_FIND_ROWS;
  int slot[SPMM_FIND_ROWS];
#pragma unroll
  for (int q = 0; q < SPMM_FIND_ROWS; q++) {          // CTA-local slots first: one global atomic per CTA, not per long row
    const long long r = base + q * blockDim.x + threadIdx.x;
    slot[q] = -1;
    if (r < rg.r1 && __ldg(pos + r + 1) - __ldg(pos + r) > cfg.thresh) slot[q] = atomicAdd(&s_cnt, 1);
  }
  __syncthreads();
  if (s_cnt == 0) return;
  if (threadIdx.x == 0) s_base = atomicAdd(counters, s_cnt);
  __syncthreads();
#pragma unroll
  for (int q = 0; q < SPMM_FIND_ROWS; q++) {
    const int i = s_base + slot[q];
    if (slot[q] >= 0 && i < cfg.nlong_max) long_rows[i] = (int)(base + q * blockDim.x + threadIdx.x);
  }
}

// The three plan kernels below run a fixed grid with grid-stride loops over the ACTUAL number of long rows (read from
// `counters[0]` on the device): their cost follows the work, not the capacity (nnz / (thresh+1) rows) the scratch is sized for.
constexpr int SPMM_PLAN_CTAS = 592;        // 148 SMs x 4

// cut[p][li] (p = 0..P) = first position of long row li whose column lies in panel p or later: one binary search per
// (row, panel boundary).
__global__ void __launch_bounds__(256)
spmm_long_cut_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const int* __restrict__ long_rows,
                     const int* __restrict__ counters, SpmmLongCfg cfg, int* __restrict__ cut) {
  const int nlong = min(__ldg(counters), cfg.nlong_max);
  const long long total = (long long)(cfg.panels + 1) * nlong;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t / nlong), li = (int)(t % nlong);
    const int r = __ldg(long_rows + li);
    const int s = __ldg(pos + r), e = __ldg(pos + r + 1);
    cut[(long long)p * cfg.nlong_max + li] = (p == 0) ? s : (p == cfg.panels) ? e : tbd::search_first_ge(crd, s, e - 1, (int)min((long long)p * cfg.panel_w, (long long)INT_MAX));
  }
}

// pair (p, li) = the nonzeros [cut[p][li], cut[p+1][li]) of long row li whose columns lie in panel p, cut into
// ceil((b-a)/cap) items.  Item slots are allocated PER PANEL: a warp (32 consecutive rows of one panel) adds up its counts
// by shuffle and takes its range with one atomicAdd on the panel's cursor (counters[8 + p]) -- no prefix sum over the
// capacity-sized pair array.  pair_loc[p][li] = first slot inside the panel, pair_cnt[p][li] = items.  The order of the
// items inside a panel depends on scheduling; the RESULT does not (every row adds its own partials in position order).
__global__ void __launch_bounds__(256)
spmm_long_alloc_kernel(const int* __restrict__ cut, int* __restrict__ counters, SpmmLongCfg cfg, int* __restrict__ pair_loc,
                       int* __restrict__ pair_cnt) {
  const int nlong = min(__ldg(counters), cfg.nlong_max);
  const int lane = threadIdx.x & 31;
  const int per = (nlong + 31) & ~31;                        // rows of a panel padded to whole warps
  const long long total = (long long)cfg.panels * per;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t / per), li = (int)(t % per);       // warp-uniform p
    int c = 0;
    if (li < nlong) {
      const long long idx = (long long)p * cfg.nlong_max + li;
      c = (__ldg(cut + idx + cfg.nlong_max) - __ldg(cut + idx) + cfg.cap - 1) / cfg.cap;
    }
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && warp_total > 0) base = atomicAdd(counters + 8 + p, warp_total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (li < nlong) {
      const long long idx = (long long)p * cfg.nlong_max + li;
      pair_loc[idx] = base + incl - c;
      pair_cnt[idx] = c;
    }
  }
}

// counters[8 + p] (items of panel p) -> counters[48 + p] = first item of panel p, counters[2] = number of items
__global__ void spmm_long_bases_kernel(int* __restrict__ counters, int panels) {
  if (threadIdx.x == 0) {
    int run = 0;
    for (int p = 0; p < panels; p++) { counters[48 + p] = run; run += counters[8 + p]; }
    counters[2] = run;
  }
}

__global__ void __launch_bounds__(256)
spmm_long_items_kernel(const int* __restrict__ cut, const int* __restrict__ pair_loc, const int* __restrict__ pair_cnt,
                       const int* __restrict__ counters, SpmmLongCfg cfg, int2* __restrict__ items) {
  const int nlong = min(__ldg(counters), cfg.nlong_max);
  const long long total = (long long)cfg.panels * nlong;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(t / nlong), li = (int)(t % nlong);
    const long long idx = (long long)p * cfg.nlong_max + li;
    const int c = __ldg(pair_cnt + idx);
    if (c == 0) continue;
    const int o = __ldg(counters + 48 + p) + __ldg(pair_loc + idx);
    const int a = __ldg(cut + idx), b = __ldg(cut + idx + cfg.nlong_max);
    for (int j = 0; j < c; j++) items[o + j] = make_int2(a + j * cfg.cap, min(a + (j + 1) * cfg.cap, b));
  }
}

// Persistent warps take items in panel order (tickets of SPMM_LONG_CHUNK items) and store one partial row sum per item.
constexpr int SPMM_LONG_CHUNK = 4;
template <typename T, int VEC, int U, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_long_kernel(const int* __restrict__ crd, const T* __restrict__ vals, const T* __restrict__ B, int K,
                 const int2* __restrict__ items, int* __restrict__ counters, T* __restrict__ partials) {
  __shared__ SpmmNz<T> stage_all[WARPS][32];
  SpmmNz<T>* stage = stage_all[threadIdx.x >> 5];
  const unsigned stride = (unsigned)K * (unsigned)sizeof(T);
  const int lane = threadIdx.x & 31;
  const int nitems = *(volatile int*)(counters + 2);
  for (;;) {
    int base = 0;
    if (lane == 0) base = atomicAdd(counters + 1, SPMM_LONG_CHUNK);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= nitems) break;
    const int end = min(base + SPMM_LONG_CHUNK, nitems);
    for (int it = base; it < end; it++) {
      const int2 m = __ldg(items + it);
      for (int c0 = 0; c0 < K; c0 += 32 * VEC) {                 // K <= 32*VEC: one pass (warp-uniform loop)
        const int col = c0 + lane * VEC;
        const bool active = col < K;
        Frag<T, VEC> acc;
#pragma unroll
        for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
        spmm_accumulate<T, VEC, U, true>(acc, crd, vals, (const char*)(B + (active ? col : 0)), stride, m.x, m.y, lane, stage);
        if (active) {
          T* dst = partials + (size_t)it * K + col;
#pragma unroll
          for (int x = 0; x < VEC; x++) dst[x] = acc.v[x];
        }
      }
    }
  }
}

// One warp per long row: its partials added in ascending panel / item (= position) order, the row of C stored once.
template <typename T, int VEC, bool COLMAJOR, bool RMAP>
__global__ void __launch_bounds__(256)
spmm_long_combine_kernel(const int* __restrict__ long_rows, const int* __restrict__ counters, const int* __restrict__ pair_loc,
                         const int* __restrict__ pair_cnt,
                         SpmmLongCfg cfg, const T* __restrict__ partials, T* __restrict__ C, int rows, int K,
                         const unsigned* __restrict__ rowmap, Fanout fo) {
  const int lane = threadIdx.x & 31;
  const int nlong = min(__ldg(counters), cfg.nlong_max);
  const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
  for (int li = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); li < nlong; li += nwarps) {   // fixed grid, actual work
  const int r = __ldg(long_rows + li);
  const size_t orow = RMAP ? (size_t)__ldg(rowmap + r) : (size_t)r;
  int o = 0, c = 0;
  if (lane < cfg.panels) {
    const long long idx = (long long)lane * cfg.nlong_max + li;
    c = __ldg(pair_cnt + idx);
    o = __ldg(counters + 48 + lane) + __ldg(pair_loc + idx);
  }
  for (int c0 = 0; c0 < K; c0 += 32 * VEC) {
    const int col = c0 + lane * VEC;
    const bool active = col < K;
    Frag<T, VEC> acc;
#pragma unroll
    for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
    for (int p = 0; p < cfg.panels; p++) {
      const int op = __shfl_sync(0xffffffffu, o, p), cp = __shfl_sync(0xffffffffu, c, p);
      for (int j = 0; j < cp; j += 4) {             // four partial rows in flight, added in item order
        Frag<T, VEC> f[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (active && j + u < cp) f[u] = load_row<T, VEC>(partials + (size_t)(op + j + u) * K + col);
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (active && j + u < cp) {
#pragma unroll
            for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + f[u].v[x];
          }
      }
    }
    if (active) store_row<T, VEC, COLMAJOR>(C, orow, col, rows, K, acc, fo);
  }
  }
}


// TACO_B200_SPMM_LONG / _PANELS / _CAP override the defaults (tuning runs).  The partial-sum scratch is sized for the worst case
// (every long row has nonzeros in every panel: 8.7 GB at C2, of which the kernel touches 0.5 GB) and kept under an eighth of the
// device memory (TACO_B200_SPMM_SCRATCH_GB overrides) by halving the panel count, then doubling the threshold.
static SpmmLongCfg spmm_long_cfg(int nnz, int cols, int K, size_t es, bool rmap) {
  static const int e_thresh = env_int("TACO_B200_SPMM_LONG", 0), e_panels = env_int("TACO_B200_SPMM_PANELS", 0),
                   e_cap = env_int("TACO_B200_SPMM_CAP", 256);
  SpmmLongCfg c;
  // (the fibers of a CSF tensor -- TTM, `rmap` -- measure 2.43 -> 2.27 ms with thresh 16 and ONE panel, profiles/r02_variants.md;
  //  not the default: it would shrink the range in which TTM is bit-identical to the reference from 128 to 16 leaves per fiber)
  (void)rmap;
  c.thresh = e_thresh > 0 ? (e_thresh < 8 ? 8 : e_thresh) : 128;
  c.panels = e_panels > 0 ? (e_panels > 32 ? 32 : e_panels) : 4;
  c.cap = e_cap < 32 ? 32 : e_cap;
  static const int e_gb = env_int("TACO_B200_SPMM_SCRATCH_GB", 0);
  const size_t budget = e_gb > 0 ? (size_t)e_gb << 30 : device_mem_total() / 8;
  for (;;) {
    c.nlong_max = nnz / (c.thresh + 1);
    c.items_max = (long long)nnz / c.cap + 1 + (long long)c.nlong_max * c.panels;
    if ((size_t)c.items_max * K * es <= budget || c.nlong_max == 0) break;
    if (c.panels > 1) c.panels /= 2;
    else if (c.thresh < (1 << 29)) c.thresh *= 2;
    else { c.nlong_max = 0; break; }          // no room even for one item per 2^29 nonzeros: every row on schedule (1)
  }
  c.panel_w = (cols + c.panels - 1) / c.panels;
  if (c.panel_w < 1) c.panel_w = 1;
  return c;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <typename T, int VEC, int U, int MINB>
static void spmm_long_go(const int* crd, const T* vals, const T* B, int K, const int2* items, int* counters, T* partials,
                         cudaStream_t st) {
  constexpr int WARPS = 8;
  // persistent grid: MINB CTAs per SM by default; TACO_B200_SPMM_LONGCTAS caps it so that the slot kernel (side by side on the
  // other stream) keeps part of every SM
  static const int cap = env_int("TACO_B200_SPMM_LONGCTAS", 0);
  const int per_sm = cap > 0 && cap < MINB ? cap : MINB;
  spmm_long_kernel<T, VEC, U, WARPS, MINB><<<num_sms() * per_sm, WARPS * 32, 0, st>>>(crd, vals, B, K, items, counters, partials);
}

// Builds the plan on the compute stream, then runs the item kernel and the combine kernel on `run_st` (the compute stream
// itself, or a side stream so that they overlap the slot kernel: the two are bound by different resources -- the panel-ordered
// gathers by the L2 -> SM path, the short rows by HBM).  *scratch_out is released by the caller after the streams have joined.
template <typename T, int VEC, bool COLMAJOR, bool RMAP>
static int spmm_long_rows(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int cols, int K, SpmmRange rg,
                          const SpmmLongCfg& cfg, const unsigned* rowmap, cudaStream_t run_st, cudaEvent_t fork, void** scratch_out,
                          Fanout fo) {
  const long long pairs = (long long)cfg.panels * cfg.nlong_max;
  // counters: [0] long rows, [1] item ticket, [2] items, [8..8+P) items per panel, [48..48+P) first item of a panel
  const size_t o_cnt = 0, o_rows = align256(sizeof(int) * 96), o_ab = o_rows + align256(sizeof(int) * (size_t)cfg.nlong_max),
               o_off = o_ab + align256(sizeof(int) * (size_t)(pairs + cfg.nlong_max)), o_pc = o_off + align256(sizeof(int) * (size_t)pairs),
               o_items = o_pc + align256(sizeof(int) * (size_t)pairs),
               o_part = o_items + align256(sizeof(int2) * (size_t)cfg.items_max),
               total = o_part + align256(sizeof(T) * (size_t)cfg.items_max * K);
  void* buf = nullptr;
  TB_TRY(scratch_alloc(&buf, total));
  *scratch_out = buf;
  char* b = (char*)buf;
  int* counters = (int*)(b + o_cnt);
  int* long_rows = (int*)(b + o_rows);
  int* cut = (int*)(b + o_ab);
  int* pair_loc = (int*)(b + o_off);
  int* pair_cnt = (int*)(b + o_pc);
  int2* items = (int2*)(b + o_items);
  T* partials = (T*)(b + o_part);
  cudaStream_t st = stream();
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(int) * 96, st);
  if (e != cudaSuccess) return fail(TACO_B200_ERR_CUDA, "spmm: memset failed: %s", cudaGetErrorString(e));
  const int nrows = rg.r1 - rg.r0;
  spmm_long_find_kernel<<<(nrows + 256 * SPMM_FIND_ROWS - 1) / (256 * SPMM_FIND_ROWS), 256, 0, st>>>(pos, rg, cfg, long_rows, counters);
  spmm_long_cut_kernel<<<SPMM_PLAN_CTAS, 256, 0, st>>>(pos, crd, long_rows, counters, cfg, cut);
  spmm_long_alloc_kernel<<<SPMM_PLAN_CTAS, 256, 0, st>>>(cut, counters, cfg, pair_loc, pair_cnt);
  spmm_long_bases_kernel<<<1, 32, 0, st>>>(counters, cfg.panels);
  spmm_long_items_kernel<<<SPMM_PLAN_CTAS, 256, 0, st>>>(cut, pair_loc, pair_cnt, counters, cfg, items);
  if (run_st != st) {
    TB_CUDA(cudaEventRecord(fork, st));
    TB_CUDA(cudaStreamWaitEvent(run_st, fork, 0));
  }
  // (gathers in flight per warp, min CTAs per SM) of the item kernel: items are long runs of one row, so per-warp
  // parallelism pays more than occupancy here (TACO_B200_SPMM_LONGVAR sweeps it)
  static const int lvar = env_int("TACO_B200_SPMM_LONGVAR", 0);
  switch (lvar) {
    case 1: spmm_long_go<T, VEC, 2, 8>(crd, vals, B, K, items, counters, partials, run_st); break;
    case 2: spmm_long_go<T, VEC, 4, 4>(crd, vals, B, K, items, counters, partials, run_st); break;
    case 3: spmm_long_go<T, VEC, 8, 4>(crd, vals, B, K, items, counters, partials, run_st); break;
    case 4: spmm_long_go<T, VEC, 8, 3>(crd, vals, B, K, items, counters, partials, run_st); break;
    default: spmm_long_go<T, VEC, 4, 6>(crd, vals, B, K, items, counters, partials, run_st); break;
  }
  spmm_long_combine_kernel<T, VEC, COLMAJOR, RMAP><<<num_sms() * 8, 256, 0, run_st>>>(
      long_rows, counters, pair_loc, pair_cnt, cfg, partials, C, rows, K, rowmap, fo);
  count_launch(7);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

// Launch variants of schedule (1): (gathers in flight per warp, warps per CTA, min CTAs per SM).  TACO_B200_SPMM_VARIANT
// selects one for tuning runs; the default is the measured best at config C2 (profiles/).
template <typename T, int VEC, bool COLMAJOR, bool RMAP, int U, int WARPS, int MINB>
static void spmm_go(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, SpmmRange rg, int nslots,
                    const int* slot_rows, int long_thresh, const unsigned* rowmap, Fanout fo, cudaStream_t st) {
  dim3 grid((nslots + WARPS - 1) / WARPS, (K + 32 * VEC - 1) / (32 * VEC));
  if (fo.n != 0 && !RMAP)
    spmm_csr_kernel<T, VEC, COLMAJOR, U, WARPS, MINB, false, true><<<grid, WARPS * 32, 0, st>>>(pos, crd, vals, B, C, rows, K, rg, nslots,
                                                                                               slot_rows, long_thresh, fo, rowmap);
  else
    spmm_csr_kernel<T, VEC, COLMAJOR, U, WARPS, MINB, RMAP, false><<<grid, WARPS * 32, 0, st>>>(pos, crd, vals, B, C, rows, K, rg, nslots,
                                                                                                slot_rows, long_thresh, Fanout(), rowmap);
}

// The whole SpMM launch sequence over a row range: long-row plan + column-panel kernels, then the slot kernel.
// `variant_env` names the environment variable that selects a launch variant of the slot kernel.
template <typename T, int VEC, bool COLMAJOR, bool RMAP>
static int spmm_launch_impl(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int cols, int K, SpmmRange rg,
                            const unsigned* rowmap, const char* prof_name) {
  const int nnz = rg.p1 - rg.p0;
  const int nslots = nnz > 0 ? (nnz + SPMM_W - 1) / SPMM_W : 1;
  void* slot_rows = nullptr;
  TB_TRY(scratch_alloc(&slot_rows, sizeof(int) * (size_t)(nslots + 1)));
  ProfScope ps(prof_name);
  // a result inside the registered fan-out window also goes to the other GPUs from inside the kernels (fused all-gather)
  const Fanout fo = RMAP ? Fanout() : result_fanout(C, sizeof(T) * (size_t)rows * K);
  const SpmmLongCfg cfg = spmm_long_cfg(nnz, cols, K, sizeof(T), RMAP);
  static const int overlap = env_int("TACO_B200_SPMM_OVERLAP", 1);
  static cudaEvent_t ev_fork = nullptr, ev_join = nullptr;          // timing-free events, created once
  if (!ev_fork) {
    TB_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    TB_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  cudaStream_t side = (overlap && cfg.nlong_max > 0) ? aux_stream(2) : stream();
  void* long_scratch = nullptr;
  int rc = TACO_B200_OK;
  if (cfg.nlong_max > 0)
    rc = spmm_long_rows<T, VEC, COLMAJOR, RMAP>(pos, crd, vals, B, C, rows, cols, K, rg, cfg, rowmap, side, ev_fork, &long_scratch, fo);
  if (rc != TACO_B200_OK) {
    if (side != stream()) { cudaEventRecord(ev_join, side); cudaStreamWaitEvent(stream(), ev_join, 0); }
    scratch_free(long_scratch); scratch_free(slot_rows);
    return rc;
  }
  spmm_slot_rows_kernel<<<(nslots + 1 + 255) / 256, 256, 0, stream()>>>(pos, rg, nslots, (int*)slot_rows);
  static const int variant = env_int(RMAP ? "TACO_B200_TTM_UNROLL" : "TACO_B200_SPMM_VARIANT", 0);
  const int* sr = (const int*)slot_rows;
  cudaStream_t st = stream();
  switch (variant) {
    case 1: spmm_go<T, VEC, COLMAJOR, RMAP, 4, 8, 4>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, cfg.thresh, rowmap, fo, st); break;
    case 2: spmm_go<T, VEC, COLMAJOR, RMAP, 1, 8, 8>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, cfg.thresh, rowmap, fo, st); break;
    case 3: spmm_go<T, VEC, COLMAJOR, RMAP, 2, 8, 6>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, cfg.thresh, rowmap, fo, st); break;
    case 4: spmm_go<T, VEC, COLMAJOR, RMAP, 4, 8, 6>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, cfg.thresh, rowmap, fo, st); break;
    default: spmm_go<T, VEC, COLMAJOR, RMAP, 2, 8, 8>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, cfg.thresh, rowmap, fo, st); break;
  }
  count_launch(2);
  if (side != stream()) {                 // join: the scratch is released (stream-ordered) only after the side stream is done
    TB_CUDA(cudaEventRecord(ev_join, side));
    TB_CUDA(cudaStreamWaitEvent(stream(), ev_join, 0));
  }
  scratch_free(long_scratch);
  scratch_free(slot_rows);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

// Row-mapped launch (csf.cu, TTM): C[rowmap[r], :] = sum_p vals[p] * B[crd[p], :] over the rows r of any (pos, crd, vals) level
// (B has `cols` rows).
int spmm_mapped(DType dt, const int* pos, const int* crd, const void* vals, const void* B, void* C, int rows, int cols, int K, int nnz,
                const unsigned* rowmap, const char* prof_name) {
  const bool a16 = (((uintptr_t)B | (uintptr_t)C) & 15) == 0;
  const SpmmRange rg{0, rows, 0, nnz};
  if (dt == DType::F64) {
    if (a16 && K % 2 == 0) return spmm_launch_impl<double, 2, false, true>(pos, crd, (const double*)vals, (const double*)B, (double*)C, rows, cols, K, rg, rowmap, prof_name);
    return spmm_launch_impl<double, 1, false, true>(pos, crd, (const double*)vals, (const double*)B, (double*)C, rows, cols, K, rg, rowmap, prof_name);
  }
  if (a16 && K % 4 == 0) return spmm_launch_impl<float, 4, false, true>(pos, crd, (const float*)vals, (const float*)B, (float*)C, rows, cols, K, rg, rowmap, prof_name);
  return spmm_launch_impl<float, 1, false, true>(pos, crd, (const float*)vals, (const float*)B, (float*)C, rows, cols, K, rg, rowmap, prof_name);
}

template <typename T>
static int spmm_launch(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int cols, int K, SpmmRange rg,
                       bool colmajor) {
  constexpr int V = 16 / sizeof(T);
  const bool vec_ok = (K % V == 0) && (((uintptr_t)B & 15) == 0) && (((uintptr_t)C & 15) == 0);
  if (colmajor) {
    if (vec_ok) return spmm_launch_impl<T, V, true, false>(pos, crd, vals, B, C, rows, cols, K, rg, nullptr, "spmm_csr");
    return spmm_launch_impl<T, 1, true, false>(pos, crd, vals, B, C, rows, cols, K, rg, nullptr, "spmm_csr");
  }
  if (vec_ok) return spmm_launch_impl<T, V, false, false>(pos, crd, vals, B, C, rows, cols, K, rg, nullptr, "spmm_csr");
  return spmm_launch_impl<T, 1, false, false>(pos, crd, vals, B, C, rows, cols, K, rg, nullptr, "spmm_csr");
}

int csr_nnz(const CsrView& A, int32_t vals_size_hint, int32_t* nnz);   // spmv.cu

static int spmm_views(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, DenseView* Cv, CsrView* Av, DenseView* Bv,
                      bool* colmajor) {
  TB_TRY(ensure_init());
  TB_TRY(view_dense(C, 2, "C", Cv));
  TB_TRY(view_csr(A, "A", Av));
  TB_TRY(view_dense(B, 2, "B", Bv));
  if (Bv->mode_order[0] != 0 || Bv->mode_order[1] != 1)
    return fail(TACO_B200_ERR_FORMAT, "spmm: B must be row-major {Dense,Dense}");
  *colmajor = (Cv->mode_order[0] == 1 && Cv->mode_order[1] == 0);
  if (!*colmajor && !(Cv->mode_order[0] == 0 && Cv->mode_order[1] == 1))
    return fail(TACO_B200_ERR_FORMAT, "spmm: bad mode ordering for C");
  if (Cv->dim[0] != Av->rows || Bv->dim[0] != Av->cols || Cv->dim[1] != Bv->dim[1])
    return fail(TACO_B200_ERR_ARG, "spmm: dimension mismatch C[%d x %d] = A[%d x %d] * B[%d x %d]", Cv->dim[0],
                Cv->dim[1], Av->rows, Av->cols, Bv->dim[0], Bv->dim[1]);
  if (Cv->dt != Av->dt || Bv->dt != Av->dt) return fail(TACO_B200_ERR_FORMAT, "spmm: mixed component types");
  return TACO_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Doubly compressed rows: A = {Compressed, Compressed} (the operand of the reference's spmmDCSRGPU test,
// /root/reference/test/tests-scheduling-eval.cpp:1309-1358, schedule scheduleSpMMNZRowsGPU :358-369).
// Level 0 stores only the rows that have nonzeros (pos0 = {0, stored rows}, crd0 = their row ids), level 1 is a CSR over
// those stored rows.  On the device the level-0 list is expanded into an ordinary CSR pos array over ALL rows
// (pos_full[i] = pos1[number of stored rows with id < i]); crd1 / vals are shared as they are.  The tuned CSR kernel
// then runs unchanged: absent rows are empty rows (zeroed by their owner, as the reference's zero-fill of C does), and
// every stored row keeps the reference's accumulation order.
// ---------------------------------------------------------------------------------------------------------
struct DcsrView { int32_t rows, cols; int32_t* pos0; int32_t* crd0; int32_t* pos1; int32_t* crd1; void* vals; DType dt; };

static int view_dcsr(const taco_tensor_t* t, const char* name, DcsrView* v) {
  if (!t) return fail(TACO_B200_ERR_ARG, "%s: NULL tensor", name);
  if (t->order != 2 || t->mode_types[0] != taco_mode_sparse || t->mode_types[1] != taco_mode_sparse ||
      t->mode_ordering[0] != 0 || t->mode_ordering[1] != 1)
    return fail(TACO_B200_ERR_FORMAT, "%s: expected DCSR ({Compressed,Compressed}, mode ordering 0,1)", name);
  v->rows = t->dimensions[0];
  v->cols = t->dimensions[1];
  if (v->rows < 0 || v->cols < 0) return fail(TACO_B200_ERR_ARG, "%s: negative dimension", name);
  if (!t->indices || !t->indices[0] || !t->indices[1]) return fail(TACO_B200_ERR_ARG, "%s: missing level arrays", name);
  v->pos0 = (int32_t*)t->indices[0][0]; v->crd0 = (int32_t*)t->indices[0][1];
  v->pos1 = (int32_t*)t->indices[1][0]; v->crd1 = (int32_t*)t->indices[1][1];
  if (!v->pos0 || !v->pos1) return fail(TACO_B200_ERR_ARG, "%s: a compressed level has no pos array", name);
  v->vals = t->vals;
  return dtype_of(t, &v->dt);
}

// One warp per stored row r (and one more for the tail): rows (crd0[r-1], crd0[r]] start at pos1[r]; the rows after the
// last stored row, and pos_full[rows], get pos1[stored].
__global__ void __launch_bounds__(256)
dcsr_expand_pos_kernel(const int* __restrict__ crd0, const int* __restrict__ pos1, int stored, int rows, int* __restrict__ pos_full) {
  const int r = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (r > stored) return;
  const int lo = r == 0 ? 0 : __ldg(crd0 + r - 1) + 1;
  const int hi = r < stored ? __ldg(crd0 + r) : rows;             // inclusive
  const int p = __ldg(pos1 + r);
  for (int i = lo + lane; i <= hi; i += 32) pos_full[i] = p;
}

static int spmm_dcsr_views(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, DenseView* Cv, DcsrView* Av, DenseView* Bv,
                           bool* colmajor) {
  TB_TRY(ensure_init());
  TB_TRY(view_dense(C, 2, "C", Cv));
  TB_TRY(view_dcsr(A, "A", Av));
  TB_TRY(view_dense(B, 2, "B", Bv));
  if (Bv->mode_order[0] != 0 || Bv->mode_order[1] != 1)
    return fail(TACO_B200_ERR_FORMAT, "spmm_dcsr: B must be row-major {Dense,Dense}");
  *colmajor = (Cv->mode_order[0] == 1 && Cv->mode_order[1] == 0);
  if (!*colmajor && !(Cv->mode_order[0] == 0 && Cv->mode_order[1] == 1))
    return fail(TACO_B200_ERR_FORMAT, "spmm_dcsr: bad mode ordering for C");
  if (Cv->dim[0] != Av->rows || Bv->dim[0] != Av->cols || Cv->dim[1] != Bv->dim[1])
    return fail(TACO_B200_ERR_ARG, "spmm_dcsr: dimension mismatch C[%d x %d] = A[%d x %d] * B[%d x %d]", Cv->dim[0],
                Cv->dim[1], Av->rows, Av->cols, Bv->dim[0], Bv->dim[1]);
  if (Cv->dt != Av->dt || Bv->dt != Av->dt) return fail(TACO_B200_ERR_FORMAT, "spmm_dcsr: mixed component types");
  return TACO_B200_OK;
}

// TACO_B200_PIPELINE_MIN_BYTES: smallest (result + nonzero) volume that takes the chunked path (default 64 MiB;
// 0 forces it, a huge value disables it)
static size_t pipeline_min_bytes() {
  const char* e = getenv("TACO_B200_PIPELINE_MIN_BYTES");
  return e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)64 << 20);
}

// Host-operand pipeline.  When A and C live in host memory the call is PCIe-bound, so it is cut into row chunks:
//   upload stream : pos, B (unless device resident), then crd/vals of chunk 0, 1, ...
//   compute stream: kernel of chunk c as soon as its nonzeros have landed
//   download stream: rows of C of chunk c as soon as its kernel is done
// Uploads of later chunks overlap the downloads of earlier ones (PCIe is full duplex), and the kernels hide under both.
template <typename T>
static int spmm_compute_pipelined(const CsrView& Av, const DenseView& Bv, const DenseView& Cv, int K, int32_t nnz) {
  const int rows = Av.rows;
  const size_t es = sizeof(T);
  cudaStream_t main = stream(), up = aux_stream(0), down = aux_stream(1);
  void *dpos = nullptr, *dcrd = nullptr, *dvals = nullptr, *dC = nullptr;
  In bin;                                           // device-resident / registered B is used in place
  const bool b_host = classify(Bv.vals) != Mem::Device && !is_resident(Bv.vals, es * (size_t)Av.cols * K);
  void* dB = nullptr;
  PipelineGuard guard;                              // scratch and events are released on every exit path
  TB_TRY(guard.alloc(&dpos, sizeof(int) * ((size_t)rows + 1)));
  TB_TRY(guard.alloc(&dcrd, sizeof(int) * (size_t)nnz));
  TB_TRY(guard.alloc(&dvals, es * (size_t)nnz));
  TB_TRY(guard.alloc(&dC, es * (size_t)rows * K));
  if (b_host) TB_TRY(guard.alloc(&dB, es * (size_t)Av.cols * K));
  else { TB_TRY(bin.acquire(Bv.vals, es * (size_t)Av.cols * K)); dB = (void*)bin.dptr; }
  cudaEvent_t ready, done_all, e_up, e_done;        // e_up / e_done are re-recorded per chunk (a wait captures the state at its call)
  TB_TRY(guard.event(&ready));
  TB_TRY(guard.event(&done_all));
  TB_TRY(guard.event(&e_up));
  TB_TRY(guard.event(&e_done));
  TB_CUDA(cudaEventRecord(ready, main));            // the pool allocations above are ordered on the compute stream
  TB_CUDA(cudaStreamWaitEvent(up, ready, 0));
  TB_CUDA(cudaStreamWaitEvent(down, ready, 0));
  TB_CUDA(cudaMemcpyAsync(dpos, Av.pos, sizeof(int) * ((size_t)rows + 1), cudaMemcpyHostToDevice, up));
  if (b_host) TB_CUDA(cudaMemcpyAsync(dB, Bv.vals, es * (size_t)Av.cols * K, cudaMemcpyHostToDevice, up));
  // chunk boundaries: balanced by (nonzeros uploaded + result bytes downloaded)
  const int nchunks = 16;
  const double w_row = (double)K * es, w_nz = 4.0 + es, total = w_row * rows + w_nz * nnz;
  int r0 = 0;
  int rc = TACO_B200_OK;
  for (int c = 0; c < nchunks && r0 < rows && rc == TACO_B200_OK; c++) {
    int r1 = rows;
    if (c < nchunks - 1) {
      const double target = total * (c + 1) / nchunks;
      int lo = r0 + 1, hi = rows;                   // smallest r1 with weight(r1) >= target
      while (lo < hi) {
        const int mid = lo + (hi - lo) / 2;
        if (w_row * mid + w_nz * Av.pos[mid] >= target) hi = mid; else lo = mid + 1;
      }
      r1 = lo;
    }
    const int p0 = Av.pos[r0], p1 = Av.pos[r1];
    if (p1 > p0) {
      TB_CUDA(cudaMemcpyAsync((int*)dcrd + p0, Av.crd + p0, sizeof(int) * (size_t)(p1 - p0), cudaMemcpyHostToDevice, up));
      TB_CUDA(cudaMemcpyAsync((T*)dvals + p0, (const T*)Av.vals + p0, es * (size_t)(p1 - p0), cudaMemcpyHostToDevice, up));
    }
    TB_CUDA(cudaEventRecord(e_up, up));
    TB_CUDA(cudaStreamWaitEvent(main, e_up, 0));
    rc = spmm_launch<T>((const int*)dpos, (const int*)dcrd, (const T*)dvals, (const T*)dB, (T*)dC, rows, Av.cols, K,
                        SpmmRange{r0, r1, p0, p1}, false);
    TB_CUDA(cudaEventRecord(e_done, main));
    TB_CUDA(cudaStreamWaitEvent(down, e_done, 0));
    TB_CUDA(cudaMemcpyAsync((T*)Cv.vals + (size_t)r0 * K, (const T*)dC + (size_t)r0 * K, es * (size_t)(r1 - r0) * K,
                            cudaMemcpyDeviceToHost, down));
    r0 = r1;
  }
  TB_CUDA(cudaEventRecord(done_all, down));
  TB_CUDA(cudaStreamWaitEvent(main, done_all, 0));  // the guard's frees (and the caller's sync) are ordered after the downloads
  TB_TRY(rc);
  TB_CUDA(cudaStreamSynchronize(main));
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_spmm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  DenseView Cv, Bv; CsrView Av; bool cm;
  TB_TRY(spmm_views(C, A, B, &Cv, &Av, &Bv, &cm));
  void* p = result_alloc(Cv.count() * dsize(Cv.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "spmm: cannot allocate result");
  C->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

int taco_b200_spmm_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  DenseView Cv, Bv; CsrView Av; bool cm;
  TB_TRY(spmm_views(C, A, B, &Cv, &Av, &Bv, &cm));
  int32_t nnz = 0;
  TB_TRY(csr_nnz(Av, A->vals_size, &nnz));
  if (nnz < 0 || nnz > INT32_MAX - 65536) return fail(TACO_B200_ERR_ARG, "spmm: bad nnz %d", nnz);
  const int K = Bv.dim[1];
  size_t es = dsize(Av.dt);
  if (!Cv.vals) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
  // host-described A and host C, large enough for chunking to pay: overlap upload / kernels / download
  if (!cm && Av.rows > 0 && K > 0 && nnz > 0 && classify(Av.pos) != Mem::Device && classify(Av.crd) != Mem::Device &&
      classify(Av.vals) != Mem::Device && classify(Cv.vals) != Mem::Device &&
      es * (size_t)Av.rows * K + (4 + es) * (size_t)nnz >= pipeline_min_bytes()) {
    if (Av.dt == DType::F32) return spmm_compute_pipelined<float>(Av, Bv, Cv, K, nnz);
    return spmm_compute_pipelined<double>(Av, Bv, Cv, K, nnz);
  }
  In pos, crd, vals, bin; Out cout;
  TB_TRY(pos.acquire(Av.pos, sizeof(int32_t) * ((size_t)Av.rows + 1)));
  TB_TRY(crd.acquire(Av.crd ? (void*)Av.crd : (void*)Av.pos, sizeof(int32_t) * (size_t)nnz));
  TB_TRY(vals.acquire(Av.vals ? Av.vals : (void*)Av.pos, es * (size_t)nnz));
  TB_TRY(bin.acquire(Bv.vals, es * (size_t)Av.cols * K));
  TB_TRY(cout.acquire(Cv.vals, es * (size_t)Av.rows * K));
  if (Av.rows > 0 && K > 0) {
    const SpmmRange all{0, Av.rows, 0, nnz};
    if (Av.dt == DType::F32)
      TB_TRY(spmm_launch<float>(pos.as<int>(), crd.as<int>(), vals.as<float>(), bin.as<float>(), cout.as<float>(),
                                Av.rows, Av.cols, K, all, cm));
    else
      TB_TRY(spmm_launch<double>(pos.as<int>(), crd.as<int>(), vals.as<double>(), bin.as<double>(), cout.as<double>(),
                                 Av.rows, Av.cols, K, all, cm));
  }
  TB_TRY(cout.commit());
  return finish_call();
}

int taco_b200_spmm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  TB_TRY(taco_b200_spmm_assemble(C, A, B));
  return taco_b200_spmm_compute(C, A, B);
}

int taco_b200_spmm_dcsr_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  DenseView Cv, Bv; DcsrView Av; bool cm;
  TB_TRY(spmm_dcsr_views(C, A, B, &Cv, &Av, &Bv, &cm));
  void* p = result_alloc(Cv.count() * dsize(Cv.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "spmm_dcsr: cannot allocate result");
  C->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

int taco_b200_spmm_dcsr_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  DenseView Cv, Bv; DcsrView Av; bool cm;
  TB_TRY(spmm_dcsr_views(C, A, B, &Cv, &Av, &Bv, &cm));
  if (!Cv.vals) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
  int32_t first = 0, stored = 0, nnz = 0;
  TB_TRY(read_i32(Av.pos0, &first));
  TB_TRY(read_i32(Av.pos0 + 1, &stored));
  if (first != 0 || stored < 0 || stored > Av.rows) return fail(TACO_B200_ERR_ARG, "spmm_dcsr: bad level-0 pos {%d, %d}", first, stored);
  if (A->vals_size > 0 && trusts_vals_size(Av.pos1)) nnz = A->vals_size;
  else TB_TRY(read_i32(Av.pos1 + stored, &nnz));
  if (nnz < 0 || nnz > INT32_MAX - 65536) return fail(TACO_B200_ERR_ARG, "spmm_dcsr: bad nnz %d", nnz);
  if (stored > 0 && !Av.crd0) return fail(TACO_B200_ERR_ARG, "spmm_dcsr: level 0 has no crd array");
  const int K = Bv.dim[1];
  const size_t es = dsize(Av.dt);
  In crd0, pos1, crd, vals, bin; Out cout;
  TB_TRY(crd0.acquire(Av.crd0 ? (void*)Av.crd0 : (void*)Av.pos1, sizeof(int32_t) * (size_t)stored));
  TB_TRY(pos1.acquire(Av.pos1, sizeof(int32_t) * ((size_t)stored + 1)));
  TB_TRY(crd.acquire(Av.crd1 ? (void*)Av.crd1 : (void*)Av.pos1, sizeof(int32_t) * (size_t)nnz));
  TB_TRY(vals.acquire(Av.vals ? Av.vals : (void*)Av.pos1, es * (size_t)nnz));
  TB_TRY(bin.acquire(Bv.vals, es * (size_t)Av.cols * K));
  TB_TRY(cout.acquire(Cv.vals, es * (size_t)Av.rows * K));
  if (Av.rows > 0 && K > 0) {
    void* pos_full = nullptr;
    TB_TRY(scratch_alloc(&pos_full, sizeof(int32_t) * ((size_t)Av.rows + 1)));
    {
      ProfScope ps("dcsr_expand_pos");
      const size_t threads = ((size_t)stored + 1) * 32;
      dcsr_expand_pos_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream()>>>(crd0.as<int>(), pos1.as<int>(), stored, Av.rows,
                                                                                     (int*)pos_full);
      count_launch(1);
      TB_CUDA(cudaGetLastError());
    }
    const SpmmRange all{0, Av.rows, 0, nnz};
    int rc;
    if (Av.dt == DType::F32)
      rc = spmm_launch<float>((const int*)pos_full, crd.as<int>(), vals.as<float>(), bin.as<float>(), cout.as<float>(), Av.rows, Av.cols, K, all, cm);
    else
      rc = spmm_launch<double>((const int*)pos_full, crd.as<int>(), vals.as<double>(), bin.as<double>(), cout.as<double>(), Av.rows, Av.cols, K, all, cm);
    scratch_free(pos_full);
    TB_TRY(rc);
  }
  TB_TRY(cout.commit());
  return finish_call();
}

int taco_b200_spmm_dcsr_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  TB_TRY(taco_b200_spmm_dcsr_assemble(C, A, B));
  return taco_b200_spmm_dcsr_compute(C, A, B);
}

}  // extern "C"
