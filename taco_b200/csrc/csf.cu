// csf.cu -- order-3 CSF kernels: MTTKRP, TTV, TTM.  B is {Compressed,Compressed,Compressed}, mode order 0,1,2.
//
//   MTTKRP  A(i,j)   = B(i,k,l) * C(k,j) * D(l,j)      replaces scheduleMTTKRPGPU (tests-scheduling-eval.cpp:327-342)
//   TTV     A(i,j)   = B(i,j,k) * c(k)                 replaces scheduleTTVGPU    (:308-325)   = SpMV over the fibers (spmv.cu)
//   TTM     A(i,j,l) = B(i,j,k) * C(k,l)               replaces scheduleTTMGPU    (:289-306)   = SpMM over the fibers (spmm.cu)
//
// The reference's GPU MTTKRP nnz-splits the leaf level, runs two block-start binary
// searches (taco_binarySearchBeforeBlock + IndirectBeforeBlock, codegen_cuda.cpp:110-141) and issues one global
// fp64 atomicAdd per (nnz, j): 6.4 G atomics at config C4, rank <= 32 only.
// Here a warp owns a mode-0 slice (= one row of A, so no atomics and no zero-fill pass for occupied rows); lanes
// own the rank columns.  The slice's leaves are contiguous in B3_crd / B_vals, so they are fetched 32 at a time
// with coalesced loads; each lane finds the fiber of its leaf with a binary search over the slice's B3_pos
// window, and the C / D factor rows are gathered 8 leaves deep (16 row loads in flight per warp).  Products are
// accumulated in leaf order with the reference association (B*C)*D -> bit-identical to the oracle.
// Unoccupied rows of A are zeroed by a memset node (they have no owner slice).
// Algorithmic bytes: nnz*(4+sizeof T) + 8*nfib + 8*nslice + sizeof T*R*(K + L + I).
#include <climits>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace tb {

constexpr int CSF_WARPS = 8;
constexpr int CSF_UNROLL = 8;

// mode: 0 = MTTKRP, 1 = TTM.  `F` = factor row length (R).
template <typename T, int MODE>
__global__ void __launch_bounds__(CSF_WARPS * 32)
csf3_rows_kernel(const int* __restrict__ B1_pos, const int* __restrict__ B1_crd, const int* __restrict__ B2_pos,
                 const int* __restrict__ B2_crd, const int* __restrict__ B3_pos, const int* __restrict__ B3_crd,
                 const T* __restrict__ Bv, const T* __restrict__ C, const T* __restrict__ D, T* __restrict__ A, int R,
                 int Kdim, int Idim) {
  const int lane = threadIdx.x & 31;
  const int nslices = __ldg(B1_pos + 1) - __ldg(B1_pos);
  const int nwork = (MODE == 0) ? nslices : __ldg(B2_pos + nslices);      // MTTKRP: slices; TTM: fibers
  for (long long wk = (long long)blockIdx.x * CSF_WARPS + (threadIdx.x >> 5); wk < nwork;
       wk += (long long)gridDim.x * CSF_WARPS) {
    int f0, f1;            // fiber range of this work item
    T* arow;
    if (MODE == 0) {
      const int iB = __ldg(B1_pos) + (int)wk;
      f0 = __ldg(B2_pos + iB); f1 = __ldg(B2_pos + iB + 1);
      const int i = __ldg(B1_crd + iB);
      arow = A + (size_t)i * R;
      // rows of A that have no slice are zeroed by the owner of the next occupied row (no separate fill pass)
      const int zlo = (wk == 0) ? 0 : __ldg(B1_crd + iB - 1) + 1;
      for (int r = zlo; r < i; r++)
        for (int j = lane; j < R; j += 32) A[(size_t)r * R + j] = T(0);
      if (wk == nwork - 1)
        for (int r = i + 1; r < Idim; r++)
          for (int j = lane; j < R; j += 32) A[(size_t)r * R + j] = T(0);
    } else {
      // TTM: one (i,j) fiber per warp; find its slice by binary search over B2_pos
      const int fb = (int)wk;
      const int iB = tbd::search_last_le(B2_pos, 0, nslices, fb);
      f0 = fb; f1 = fb + 1;
      arow = A + ((size_t)__ldg(B1_crd + iB) * Kdim + __ldg(B2_crd + fb)) * R;
    }
    const int l0 = __ldg(B3_pos + f0), l1 = __ldg(B3_pos + f1);
    for (int j0 = 0; j0 < R; j0 += 32) {
      const int j = j0 + lane;
      const bool active = j < R;
      T acc = T(0);
      for (int pb = l0; pb < l1; pb += 32) {
        const int cnt = min(32, l1 - pb);
        int my_l = 0, my_k = 0;
        T my_v = T(0);
        if (lane < cnt) {
          const int p = pb + lane;
          my_l = tbd::ldg_stream_i32(B3_crd + p);
          my_v = __ldg(Bv + p);
          if (MODE == 0) my_k = __ldg(B2_crd + tbd::search_last_le(B3_pos, f0, f1 - 1, p));
        }
        for (int q0 = 0; q0 < cnt; q0 += CSF_UNROLL) {
          T cv[CSF_UNROLL], dv[CSF_UNROLL];
#pragma unroll
          for (int u = 0; u < CSF_UNROLL; u++) {
            if (q0 + u < cnt) {
              const int l = __shfl_sync(0xffffffffu, my_l, q0 + u);
              if (MODE == 0) {
                const int k = __shfl_sync(0xffffffffu, my_k, q0 + u);
                if (active) { cv[u] = __ldg(C + (size_t)k * R + j); dv[u] = __ldg(D + (size_t)l * R + j); }
              } else {
                if (active) cv[u] = __ldg(C + (size_t)l * R + j);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < CSF_UNROLL; u++) {
            if (q0 + u < cnt) {
              const T v = __shfl_sync(0xffffffffu, my_v, q0 + u);
              if (active) {
                if (MODE == 0) acc = acc + (v * cv[u]) * dv[u];
                else acc = acc + v * cv[u];
              }
            }
          }
        }
      }
      if (active) arow[j] = acc;
    }
  }
}

// TTV: A(i,j) = sum_k B(i,j,k) c(k).  One thread group of 8 lanes per fiber; fibers are short at the target shapes.
template <typename T>
__global__ void __launch_bounds__(256)
csf3_ttv_kernel(const int* __restrict__ B1_pos, const int* __restrict__ B1_crd, const int* __restrict__ B2_pos,
                const int* __restrict__ B2_crd, const int* __restrict__ B3_pos, const int* __restrict__ B3_crd,
                const T* __restrict__ Bv, const T* __restrict__ c, T* __restrict__ A, int Kdim) {
  const int nslices = __ldg(B1_pos + 1) - __ldg(B1_pos);
  const int nfib = __ldg(B2_pos + nslices);
  for (long long fb = (long long)blockIdx.x * blockDim.x + threadIdx.x; fb < nfib;
       fb += (long long)gridDim.x * blockDim.x) {
    const int iB = tbd::search_last_le(B2_pos, 0, nslices, (int)fb);
    T acc = T(0);
    for (int p = __ldg(B3_pos + fb); p < __ldg(B3_pos + fb + 1); p++) acc += __ldg(Bv + p) * __ldg(c + __ldg(B3_crd + p));
    A[(size_t)__ldg(B1_crd + iB) * Kdim + __ldg(B2_crd + fb)] = acc;
  }
}

// =========================================================================================================
// MTTKRP: nnz-balanced slice-aligned slots
// =========================================================================================================
// The leaf level is cut into slots of MK_W leaves; slot w (one warp) OWNS the mode-0 slices whose first leaf lies in
// [w*W,(w+1)*W) and computes their rows of A completely in registers (lane <-> rank column), storing each row once:
// no atomics, no zero-fill pass, summation in leaf order with the reference association (B*C)*D => bit-identical to
// the reference's C kernel.  Only slices longer than MK_LONG leaves are split across the slots they span; their pieces are
// added into the row IN SLOT ORDER (the owner stores its piece, every later piece waits for its predecessor's flag, adds,
// and raises its own), so the result does not depend on scheduling: run-to-run deterministic, no atomics on values.
// Slot numbers are handed out by a ticket per CTA, so a piece's predecessor has always started -- the chain cannot deadlock
// whatever order the hardware dispatches CTAs in.  Per 32 leaves: one coalesced load of B3_crd / B_vals, the fiber of each
// leaf from a guessed position (exact when fibers are singletons) or a binary search in the slice's B3_pos window,
// (k, l, val) staged in shared memory and read back as one 16-byte broadcast per leaf.  C(k,:) is requested per leaf, but
// consecutive leaves of a fiber share k and the repeats are L1 hits: holding the row in a register across a fiber (the
// reference's `w` workspace idea) was measured and is SLOWER -- 6.70 ms against 6.17 ms on the 12.5-leaves-per-fiber tensor
// of bench.py (mttkrp_fibers), no change at C4 -- because the conditional load stops ptxas from batching the gathers.
constexpr int MK_W = 64;
constexpr int MK_LONG = 512;

template <typename T> struct MkLeaf { int k, l; T v; };
template <> struct __align__(16) MkLeaf<double> { int k, l; double v; };
template <> struct __align__(16) MkLeaf<float> { int k, l; float v; int pad; };

template <typename T>
__device__ __forceinline__ T ld_keep(const T* p, uint64_t keep) {
  T r;
  if constexpr (sizeof(T) == 8) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(keep));
  else asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(keep));
  return r;
}
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v, uint64_t strm) {
  if constexpr (sizeof(T) == 8) asm volatile("st.global.L1::no_allocate.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(strm) : "memory");
  else asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(strm) : "memory");
}

// slot_slices[w] = first slice s (0..nslices) whose first leaf B3_pos[B2_pos[s]] is >= w*W; slot_slices[nslots] = nslices.
__global__ void mttkrp_slot_slices_kernel(const int* __restrict__ B1_pos, const int* __restrict__ B2_pos,
                                          const int* __restrict__ B3_pos, int nslots, int* __restrict__ slot_slices) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nslots) return;
  const int s_base = __ldg(B1_pos), nslices = __ldg(B1_pos + 1) - s_base;
  if (w == nslots) { slot_slices[w] = nslices; return; }
  const int target = w * MK_W;
  int lo = 0, end = nslices + 1;                 // smallest s in [0, nslices] with leafstart(s) >= target
  while (lo < end) {
    const int mid = lo + ((end - lo) >> 1);
    if (__ldg(B3_pos + __ldg(B2_pos + s_base + mid)) >= target) end = mid; else lo = mid + 1;
  }
  slot_slices[w] = lo;
}

struct MkSlice { int f0, f1, l0, l1, row, zlo; };

// result store with its fan-out (fused all-gather of A's rows, common.cuh): local only, one store through the NVLink
// multicast mapping, or the local store plus one store into each peer GPU's copy
template <typename T>
__device__ __forceinline__ void mk_store(T* dst, T v, const Fanout& fo) {
  if (fo.n >= 0) {
    *dst = v;
    for (int i = 0; i < fo.n; i++) *(T*)((char*)dst + fo.d[i]) = v;
  } else if constexpr (sizeof(T) == 4) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"((char*)dst + fo.d[0]), "f"(v) : "memory");
  else asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"((char*)dst + fo.d[0]), "d"(v) : "memory");
}   // fibers, leaves, row of A, first row to zero before `row`

template <typename T, int U, int WARPS, int MINB, bool SINGLE, bool MC>
__global__ void __launch_bounds__(WARPS * 32, MINB)
mttkrp_csf_kernel(const int* __restrict__ B1_pos, const int* __restrict__ B1_crd, const int* __restrict__ B2_pos,
                  const int* __restrict__ B2_crd, const int* __restrict__ B3_pos, const int* __restrict__ B3_crd,
                  const T* __restrict__ Bv, const T* __restrict__ C, const T* __restrict__ D, T* __restrict__ A, int R,
                  int Idim, int nnz, int nslots, const int* __restrict__ slot_slices, int* __restrict__ chain, bool ticketed, Fanout fo_arg) {
  Fanout fo;                                      // a template flag: the single-GPU instantiation keeps the offsets out of its registers
  if constexpr (MC) fo = fo_arg;
  __shared__ MkLeaf<T> stage_all[WARPS][32];
  __shared__ MkSlice meta_all[WARPS][32];
  __shared__ int s_ticket;
  const int lane = threadIdx.x & 31;
  // chain[0] = CTA ticket, chain[1 + w] = "the pieces of a hub slice up to slot w are in the row" flag (zeroed per launch)
  if (ticketed) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(chain, 1);
    __syncthreads();
  }
  const int w = (ticketed ? s_ticket : (int)blockIdx.x) * WARPS + (threadIdx.x >> 5);
  int* const flags = chain + 1;
  if (w >= nslots) return;
  MkLeaf<T>* stage = stage_all[threadIdx.x >> 5];
  MkSlice* meta = meta_all[threadIdx.x >> 5];
  const int lo = w * MK_W, hi = min(lo + MK_W, nnz);
  if (lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(B3_crd + lo + lane * 32));
  else if (lane < 2 + (int)(2 * sizeof(T) / 4)) asm volatile("prefetch.global.L2 [%0];" ::"l"(Bv + lo + (lane - 2) * (128 / (int)sizeof(T))));
  const int s_base = __ldg(B1_pos), nslices = __ldg(B1_pos + 1) - s_base;
  const int S0 = __ldg(slot_slices + w), S1 = __ldg(slot_slices + w + 1);

  // group -1 (optional): the piece [lo, min(hi, end)) of a hub slice that started in an earlier slot (added in slot order);
  // groups 0..: the slices this slot owns, 32 at a time (metadata parked in shared memory, one broadcast read per slice)
  bool tail = false;
  int tail_end = 0;                        // where that hub slice really ends
  if (S0 > 0 && lo < nnz) {
    const int f0 = __ldg(B2_pos + s_base + S0 - 1), f1 = __ldg(B2_pos + s_base + S0);
    const int l0 = __ldg(B3_pos + f0), l1 = __ldg(B3_pos + f1);
    tail = l1 > lo && l1 - l0 > MK_LONG;
    tail_end = l1;
    if (tail && lane == 0) meta[0] = MkSlice{f0, f1, lo, min(hi, l1), __ldg(B1_crd + s_base + S0 - 1), -1};
  }
  for (int sb = tail ? S0 - 32 : S0; sb < S1; sb += 32) {
    const bool is_tail = sb < S0;
    int nwork = 1;
    if (!is_tail) {
      nwork = min(32, S1 - sb);
      __syncwarp();
      if (lane < nwork) {
        const int sl = sb + lane, s = s_base + sl;
        MkSlice m;
        m.f0 = __ldg(B2_pos + s); m.f1 = __ldg(B2_pos + s + 1);
        m.l0 = __ldg(B3_pos + m.f0); m.l1 = __ldg(B3_pos + m.f1);
        m.row = __ldg(B1_crd + s);
        // rows of A without a slice are zeroed by the owner of the next occupied row
        m.zlo = (sl == 0) ? 0 : __ldg(B1_crd + s - 1) + 1;
        meta[lane] = m;
      }
    }
    __syncwarp();
    for (int h = 0; h < nwork; h++) {
      const MkSlice m = meta[h];
      const bool hub = is_tail || m.l1 - m.l0 > MK_LONG;
      const int l1 = hub ? min(hi, m.l1) : m.l1;
      if (!is_tail) {
        for (int r = m.zlo; r < m.row; r++)
          for (int j = lane; j < R; j += 32) mk_store<T>(A + (size_t)r * R + j, T(0), fo);
        if (sb + h == nslices - 1)                       // the last slice also owns the rows after it
          for (int r = m.row + 1; r < Idim; r++)
            for (int j = lane; j < R; j += 32) mk_store<T>(A + (size_t)r * R + j, T(0), fo);
      }
      for (int j0 = 0; j0 < (SINGLE ? 1 : R); j0 += 32) {      // SINGLE: R <= 32, one pass over the slice's leaves
        const bool active = j0 + lane < R;
        const T* Cj = C + (active ? j0 + lane : 0);
        const T* Dj = D + (active ? j0 + lane : 0);
        T acc = T(0);
        for (int pb = m.l0; pb < l1; pb += 32) {
          const int cnt = min(32, l1 - pb);
          if (lane < cnt) {
            const int p = pb + lane;
            MkLeaf<T> e;
            e.l = tbd::ldg_stream_i32(B3_crd + p);
            e.v = __ldg(Bv + p);
            int f = min(m.f0 + (p - m.l0), m.f1 - 1);         // exact when every fiber of the slice is a singleton
            if (!(__ldg(B3_pos + f) <= p && __ldg(B3_pos + f + 1) > p)) f = tbd::search_last_le(B3_pos, m.f0, m.f1 - 1, p);
            e.k = __ldg(B2_crd + f);
            stage[lane] = e;
          }
          __syncwarp();
          int q = 0;
          for (; q + U <= cnt; q += U) {
            T cv[U], dv[U], vv[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              const MkLeaf<T> e = stage[q + u];
              vv[u] = e.v;
              cv[u] = __ldg(Cj + (size_t)e.k * R);       // consecutive leaves of a fiber share k: L1 serves the repeats
              dv[u] = __ldg(Dj + (size_t)e.l * R);
            }
#pragma unroll
            for (int u = 0; u < U; u++) acc = acc + (vv[u] * cv[u]) * dv[u];
          }
          for (; q < cnt; q++) {
            const MkLeaf<T> e = stage[q];
            acc = acc + (e.v * __ldg(Cj + (size_t)e.k * R)) * __ldg(Dj + (size_t)e.l * R);
          }
          __syncwarp();
        }
        if (is_tail && j0 == 0) {                          // a later piece of a hub slice: wait for the pieces before it
          if (lane == 0) while (atomicAdd(flags + w - 1, 0) == 0) __nanosleep(64);
          __syncwarp();
          __threadfence();
        }
        if (active) {
          T* dst = A + (size_t)m.row * R + j0 + lane;
          // pieces of a hub slice are added in slot order (deterministic) in LOCAL memory; only the piece that completes
          // the row -- and every whole slice -- goes out through the multicast mapping, if there is one
          if (is_tail) {
            const T sum = __ldcg(dst) + acc;
            if (tail_end <= hi) mk_store<T>(dst, sum, fo); else *dst = sum;
          } else if (hub) *dst = acc;
          else mk_store<T>(dst, acc, fo);
        }
      }
      if (hub && (is_tail ? tail_end > hi : true)) {       // more pieces follow in the next slot: publish this one
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicExch(flags + w, 1);
      }
    }
  }
}

// device-resident CSF: nslices = pos1[1] - pos1[0], nfib = pos2[nslices], nnz = pos3[nfib] (a negative size stops the chain)
__global__ void csf3_level_sizes_kernel(const int* __restrict__ pos1, const int* __restrict__ pos2, const int* __restrict__ pos3,
                                        int* __restrict__ out) {
  const int nslices = pos1[1] - pos1[0];
  const int nfib = nslices >= 0 ? pos2[nslices] : -1;
  const int nnz = nfib >= 0 ? pos3[nfib] : -1;
  out[0] = nslices; out[1] = nfib; out[2] = nnz;
}

struct CsfCall {
  Csf3View B; DType dt; int32_t nslices, nfib, nnz;
  In p1, c1, p2, c2, p3, c3, vals;
  bool host_described = false;       // every level array and the values live in host memory (read directly by the host)
  bool uploaded = false;
  int upload();                      // make the level arrays visible to the device (staging copies for host arrays)
};

static int csf_prepare(taco_tensor_t* Bt, CsfCall* cc) {
  TB_TRY(view_csf3(Bt, "B", &cc->B));
  cc->dt = cc->B.dt;
  // level sizes: pos1[1], pos2[nslices], pos3[nfib].  Host arrays are read directly; when all three pos arrays are device
  // memory the chain is followed by one single-thread kernel and read back with ONE synchronisation (it was one per level)
  if (classify(cc->B.pos[0]) == Mem::Device && classify(cc->B.pos[1]) == Mem::Device && classify(cc->B.pos[2]) == Mem::Device) {
    std::lock_guard<std::mutex> lk(small_scratch_mutex());
    int* d = (int*)small_scratch();
    if (!d) return fail(TACO_B200_ERR_ALLOC, "no device scratch");
    csf3_level_sizes_kernel<<<1, 1, 0, stream()>>>(cc->B.pos[0], cc->B.pos[1], cc->B.pos[2], d);
    int32_t h[3];
    TB_TRY(read_back(h, d, sizeof(h)));
    cc->nslices = h[0]; cc->nfib = h[1]; cc->nnz = h[2];
  } else {
    int32_t p10 = 0;
    TB_TRY(read_i32(cc->B.pos[0], &p10));
    TB_TRY(read_i32(cc->B.pos[0] + 1, &cc->nslices));
    cc->nslices -= p10;
    TB_TRY(read_i32(cc->B.pos[1] + cc->nslices, &cc->nfib));
    if (Bt->vals_size > 0 && trusts_vals_size(cc->B.pos[2])) cc->nnz = Bt->vals_size;
    else TB_TRY(read_i32(cc->B.pos[2] + cc->nfib, &cc->nnz));
  }
  if (cc->nslices < 0 || cc->nfib < 0 || cc->nnz < 0) return fail(TACO_B200_ERR_ARG, "csf: corrupt pos arrays");
  cc->host_described = true;
  for (int l = 0; l < 3; l++) {
    if (classify(cc->B.pos[l]) == Mem::Device) cc->host_described = false;
    if (cc->B.crd[l] && classify(cc->B.crd[l]) == Mem::Device) cc->host_described = false;
  }
  if (!cc->B.vals || classify(cc->B.vals) == Mem::Device) cc->host_described = false;
  if (cc->host_described && is_resident(cc->B.vals, dsize(cc->dt) * (size_t)cc->nnz)) cc->host_described = false;   // already in HBM
  return TACO_B200_OK;
}

int CsfCall::upload() {
  if (uploaded) return TACO_B200_OK;
  uploaded = true;
  CsfCall* cc = this;
  void* dummy = (void*)cc->B.pos[0];
  TB_TRY(cc->p1.acquire(cc->B.pos[0], sizeof(int32_t) * 2));
  TB_TRY(cc->c1.acquire(cc->B.crd[0] ? (void*)cc->B.crd[0] : dummy, sizeof(int32_t) * (size_t)cc->nslices));
  TB_TRY(cc->p2.acquire(cc->B.pos[1], sizeof(int32_t) * ((size_t)cc->nslices + 1)));
  TB_TRY(cc->c2.acquire(cc->B.crd[1] ? (void*)cc->B.crd[1] : dummy, sizeof(int32_t) * (size_t)cc->nfib));
  TB_TRY(cc->p3.acquire(cc->B.pos[2], sizeof(int32_t) * ((size_t)cc->nfib + 1)));
  TB_TRY(cc->c3.acquire(cc->B.crd[2] ? (void*)cc->B.crd[2] : dummy, sizeof(int32_t) * (size_t)cc->nnz));
  TB_TRY(cc->vals.acquire(cc->B.vals ? cc->B.vals : dummy, dsize(cc->dt) * (size_t)cc->nnz));
  return TACO_B200_OK;
}

static int dense_assemble(taco_tensor_t* A, int order, const char* what) {
  DenseView Av;
  TB_TRY(ensure_init());
  TB_TRY(view_dense(A, order, "A", &Av));
  void* p = result_alloc(Av.count() * dsize(Av.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "%s: cannot allocate result", what);
  A->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

template <typename T, int U, int WARPS, int MINB>
static void mttkrp_go(CsfCall& cc, const T* C, const T* D, T* A, int R, int nslots, const int* slot_slices, int* chain, Fanout fo) {
  const dim3 grid((nslots + WARPS - 1) / WARPS);
  // TACO_B200_MTTKRP_NOTICKET=1: slot numbers from blockIdx instead of the ticket (A/B measurements only: the ordered
  // hand-over of hub slices then relies on in-order CTA dispatch)
  static const bool ticketed = getenv("TACO_B200_MTTKRP_NOTICKET") == nullptr;
  // The single-pass specialisation (R <= 32) spills less but measures SLOWER at C4 (23.1 vs 16.7 ms, same box, A/B):
  // ptxas unrolls the leaf loop of the general version four deep, which is what keeps HBM at 98 % of its peak.
  static const bool single = getenv("TACO_B200_MTTKRP_SINGLE") != nullptr;
#define TB_MK_GO(SINGLE, MC)                                                                                                   \
  mttkrp_csf_kernel<T, U, WARPS, MINB, SINGLE, MC><<<grid, WARPS * 32, 0, stream()>>>(                                            \
      cc.p1.as<int>(), cc.c1.as<int>(), cc.p2.as<int>(), cc.c2.as<int>(), cc.p3.as<int>(), cc.c3.as<int>(), cc.vals.as<T>(), C, D, A, R, \
      cc.B.dim[0], cc.nnz, nslots, slot_slices, chain, ticketed, fo)
  if (R <= 32 && single) { if (fo.n) TB_MK_GO(true, true); else TB_MK_GO(true, false); }
  else { if (fo.n) TB_MK_GO(false, true); else TB_MK_GO(false, false); }
#undef TB_MK_GO
}

template <typename T>
static int mttkrp_launch(CsfCall& cc, const T* C, const T* D, T* A, size_t a_count, int R) {
  if (cc.nslices == 0 || R == 0 || cc.nnz == 0) {
    // no slice owns any row (or every slice is empty): the result is all zeros
    TB_CUDA(cudaMemsetAsync(A, 0, a_count * sizeof(T), stream()));
    count_launch(1);
    return TACO_B200_OK;
  }
  if (cc.nnz > INT32_MAX - 65536) return fail(TACO_B200_ERR_ARG, "mttkrp: nnz too close to the int32 limit");
  const int nslots = (cc.nnz + MK_W - 1) / MK_W;
  void* slot_slices = nullptr;
  TB_TRY(scratch_alloc(&slot_slices, sizeof(int) * (size_t)(nslots + 1)));
  mttkrp_slot_slices_kernel<<<(nslots + 1 + 255) / 256, 256, 0, stream()>>>(cc.p1.as<int>(), cc.p2.as<int>(), cc.p3.as<int>(), nslots,
                                                                           (int*)slot_slices);
  void* chain = nullptr;                      // CTA ticket + one flag per slot (only hub slices ever touch the flags)
  if (scratch_alloc(&chain, sizeof(int) * ((size_t)nslots + 1)) != TACO_B200_OK) { scratch_free(slot_slices); return TACO_B200_ERR_ALLOC; }
  if (cudaMemsetAsync(chain, 0, sizeof(int) * ((size_t)nslots + 1), stream()) != cudaSuccess) {
    scratch_free(chain); scratch_free(slot_slices);
    return fail(TACO_B200_ERR_CUDA, "mttkrp: memset failed");
  }
  static const int variant = getenv("TACO_B200_MTTKRP_VARIANT") ? atoi(getenv("TACO_B200_MTTKRP_VARIANT")) : 0;
  const Fanout fo = result_fanout(A, a_count * sizeof(T));     // result inside the registered fan-out window?
  {
    ProfScope ps("mttkrp_csf");
    const int* ss = (const int*)slot_slices;
    switch (variant) {
      case 1: mttkrp_go<T, 2, 8, 5>(cc, C, D, A, R, nslots, ss, (int*)chain, fo); break;
      case 2: mttkrp_go<T, 2, 8, 6>(cc, C, D, A, R, nslots, ss, (int*)chain, fo); break;
      case 3: mttkrp_go<T, 1, 8, 6>(cc, C, D, A, R, nslots, ss, (int*)chain, fo); break;
      case 4: mttkrp_go<T, 2, 8, 8>(cc, C, D, A, R, nslots, ss, (int*)chain, fo); break;
      case 5: mttkrp_go<T, 1, 16, 4>(cc, C, D, A, R, nslots, ss, (int*)chain, fo); break;
      default: mttkrp_go<T, 1, 8, 8>(cc, C, D, A, R, nslots, ss, (int*)chain, fo); break;
    }
  }
  count_launch(2);
  scratch_free(chain);
  scratch_free(slot_slices);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Host-operand pipeline for MTTKRP (the e2e path: B and A in host memory).  Mode-0 slices are independent, so the call is
// cut into slice chunks, each a self-contained sub-problem over its own row range of A:
//   upload stream : level arrays and values of chunk c (C and D were uploaded before the loop)
//   compute stream: rebase the chunk's pos / row ids to chunk-local numbering, then the ordinary slot kernel on it
//   download stream: the chunk's rows of A
// Uploads of later chunks overlap kernels and downloads of earlier ones (PCIe is full duplex).  Values are bit-identical
// to the unpipelined call: the kernel and the per-slice order are the same.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) csf_rebase_kernel(int* __restrict__ a, long long n, int off) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] -= off;
}

static size_t csf_pipeline_min_bytes() {
  const char* e = getenv("TACO_B200_PIPELINE_MIN_BYTES");
  return e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)64 << 20);
}

template <typename T>
static int mttkrp_compute_pipelined(CsfCall& cc, const T* C, const T* D, T* hostA, int R) {
  const size_t es = sizeof(T);
  const int I = cc.B.dim[0], ns = cc.nslices, nf = cc.nfib, nz = cc.nnz;
  const int32_t *h_c1 = cc.B.crd[0], *h_p2 = cc.B.pos[1], *h_c2 = cc.B.crd[1], *h_p3 = cc.B.pos[2], *h_c3 = cc.B.crd[2];
  const T* h_vals = (const T*)cc.B.vals;
  const int s_base = cc.B.pos[0][0];
  cudaStream_t main = stream(), up = aux_stream(0), down = aux_stream(1);
  const int nchunks = 16;
  void *d_p1 = nullptr, *d_c1 = nullptr, *d_p2 = nullptr, *d_c2 = nullptr, *d_p3 = nullptr, *d_c3 = nullptr, *d_vals = nullptr, *dA = nullptr;
  PipelineGuard guard;                              // scratch and events are released on every exit path
  TB_TRY(guard.alloc(&d_p1, sizeof(int) * 2 * nchunks));
  TB_TRY(guard.alloc(&d_c1, sizeof(int) * (size_t)ns));
  TB_TRY(guard.alloc(&d_p2, sizeof(int) * ((size_t)ns + nchunks)));       // chunk c lives at offset s0 + c (own closing entry)
  TB_TRY(guard.alloc(&d_c2, sizeof(int) * (size_t)nf));
  TB_TRY(guard.alloc(&d_p3, sizeof(int) * ((size_t)nf + nchunks)));
  TB_TRY(guard.alloc(&d_c3, sizeof(int) * (size_t)nz));
  TB_TRY(guard.alloc(&d_vals, es * (size_t)nz));
  TB_TRY(guard.alloc(&dA, es * (size_t)I * R));
  cudaEvent_t ready, done_all, e_up, e_done;        // e_up / e_done are re-recorded per chunk
  TB_TRY(guard.event(&ready));
  TB_TRY(guard.event(&done_all));
  TB_TRY(guard.event(&e_up));
  TB_TRY(guard.event(&e_done));
  TB_CUDA(cudaEventRecord(ready, main));            // pool allocations and the C / D uploads are ordered on the compute stream
  TB_CUDA(cudaStreamWaitEvent(up, ready, 0));
  TB_CUDA(cudaStreamWaitEvent(down, ready, 0));
  // chunk boundaries balanced by transferred bytes: leaves and fibers up, rows of A down
  auto weight = [&](int s) -> double {               // bytes moved for slices [0, s)
    const int f = h_p2[s_base + s];
    const double rows = s < ns ? (double)h_c1[s_base + s] : (double)I;
    return (4.0 + es) * h_p3[f] + 8.0 * f + (double)R * es * rows;
  };
  const double total = weight(ns);
  std::vector<int> hp1(2 * nchunks, 0);
  int s0 = 0, rc = TACO_B200_OK;
  std::vector<int> bounds;
  bounds.push_back(0);
  for (int c = 0; c < nchunks - 1; c++) {
    const double target = total * (c + 1) / nchunks;
    int lo = bounds.back(), hi = ns;
    while (lo < hi) { const int mid = lo + (hi - lo) / 2; if (weight(mid) >= target) hi = mid; else lo = mid + 1; }
    bounds.push_back(lo);
  }
  bounds.push_back(ns);
  for (int c = 0; c < nchunks; c++) hp1[2 * c + 1] = bounds[c + 1] - bounds[c];
  TB_CUDA(cudaMemcpyAsync(d_p1, hp1.data(), sizeof(int) * 2 * nchunks, cudaMemcpyHostToDevice, up));
  TB_CUDA(cudaStreamSynchronize(up));                // hp1 is a local: its copy must not outlive it (tiny)
  for (int c = 0; c < nchunks && rc == TACO_B200_OK; c++) {
    s0 = bounds[c];
    const int s1 = bounds[c + 1];
    const int row0 = c == 0 ? 0 : h_c1[s_base + s0];
    const int row1 = s1 < ns ? h_c1[s_base + s1] : I;
    if (s1 == s0 && row1 == row0) continue;
    const int f0 = h_p2[s_base + s0], f1 = h_p2[s_base + s1], l0 = h_p3[f0], l1 = h_p3[f1];
    int* c1 = (int*)d_c1 + s0; int* p2 = (int*)d_p2 + s0 + c; int* c2 = (int*)d_c2 + f0; int* p3 = (int*)d_p3 + f0 + c;
    int* c3 = (int*)d_c3 + l0; T* vv = (T*)d_vals + l0;
    if (s1 > s0) {
      TB_CUDA(cudaMemcpyAsync(c1, h_c1 + s_base + s0, sizeof(int) * (size_t)(s1 - s0), cudaMemcpyHostToDevice, up));
      TB_CUDA(cudaMemcpyAsync(p2, h_p2 + s_base + s0, sizeof(int) * (size_t)(s1 - s0 + 1), cudaMemcpyHostToDevice, up));
      TB_CUDA(cudaMemcpyAsync(p3, h_p3 + f0, sizeof(int) * (size_t)(f1 - f0 + 1), cudaMemcpyHostToDevice, up));
      if (f1 > f0) TB_CUDA(cudaMemcpyAsync(c2, h_c2 + f0, sizeof(int) * (size_t)(f1 - f0), cudaMemcpyHostToDevice, up));
      if (l1 > l0) {
        TB_CUDA(cudaMemcpyAsync(c3, h_c3 + l0, sizeof(int) * (size_t)(l1 - l0), cudaMemcpyHostToDevice, up));
        TB_CUDA(cudaMemcpyAsync(vv, h_vals + l0, es * (size_t)(l1 - l0), cudaMemcpyHostToDevice, up));
      }
    }
    TB_CUDA(cudaEventRecord(e_up, up));
    TB_CUDA(cudaStreamWaitEvent(main, e_up, 0));
    if (s1 > s0) {
      csf_rebase_kernel<<<(unsigned)((s1 - s0 + 255) / 256), 256, 0, main>>>(c1, s1 - s0, row0);
      csf_rebase_kernel<<<(unsigned)((s1 - s0 + 1 + 255) / 256), 256, 0, main>>>(p2, s1 - s0 + 1, f0);
      csf_rebase_kernel<<<(unsigned)((f1 - f0 + 1 + 255) / 256), 256, 0, main>>>(p3, (long long)f1 - f0 + 1, l0);
      count_launch(3);
    }
    CsfCall sub;
    sub.B = cc.B; sub.B.dim[0] = row1 - row0; sub.dt = cc.dt;
    sub.nslices = s1 - s0; sub.nfib = f1 - f0; sub.nnz = l1 - l0;
    sub.p1.dptr = (int*)d_p1 + 2 * c; sub.c1.dptr = c1; sub.p2.dptr = p2; sub.c2.dptr = c2; sub.p3.dptr = p3; sub.c3.dptr = c3;
    sub.vals.dptr = vv;
    T* Achunk = (T*)dA + (size_t)row0 * R;
    rc = mttkrp_launch<T>(sub, C, D, Achunk, (size_t)(row1 - row0) * R, R);
    TB_CUDA(cudaEventRecord(e_done, main));
    TB_CUDA(cudaStreamWaitEvent(down, e_done, 0));
    if (row1 > row0)
      TB_CUDA(cudaMemcpyAsync(hostA + (size_t)row0 * R, Achunk, es * (size_t)(row1 - row0) * R, cudaMemcpyDeviceToHost, down));
  }
  TB_CUDA(cudaEventRecord(done_all, down));
  TB_CUDA(cudaStreamWaitEvent(main, done_all, 0));    // the guard's frees are ordered after the downloads
  TB_TRY(rc);
  TB_CUDA(cudaStreamSynchronize(main));
  return TACO_B200_OK;
}

// spmv.cu / spmm.cu
int spmv_mapped(DType dt, const int* pos, const int* crd, const void* vals, const void* x, void* y, int rows, int nnz,
                const unsigned* ymap, const char* prof_name);
int spmm_mapped(DType dt, const int* pos, const int* crd, const void* vals, const void* B, void* C, int rows, int cols, int K, int nnz,
                const unsigned* rowmap, const char* prof_name);

// fiber_cell[f] = B1_crd[s] * Kdim + B2_crd[f] for every fiber f of slice s: where the fiber's result lives in the dense
// (i,j) plane of A.  One warp per slice, coalesced over its fibers.
__global__ void __launch_bounds__(256)
csf3_fiber_cells_kernel(const int* __restrict__ B1_pos, const int* __restrict__ B1_crd, const int* __restrict__ B2_pos,
                        const int* __restrict__ B2_crd, unsigned Kdim, unsigned* __restrict__ cell) {
  const int nslices = __ldg(B1_pos + 1) - __ldg(B1_pos);
  const int lane = threadIdx.x & 31;
  for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < nslices; s += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int iB = __ldg(B1_pos) + (int)s;
    const unsigned base = (unsigned)__ldg(B1_crd + iB) * Kdim;
    const int f1 = __ldg(B2_pos + iB + 1);
    for (int f = __ldg(B2_pos + iB) + lane; f < f1; f += 32) cell[f] = base + (unsigned)__ldg(B2_crd + f);
  }
}

static int fiber_cells(CsfCall& cc, int Kdim, void** cell) {
  TB_TRY(scratch_alloc(cell, sizeof(unsigned) * (size_t)cc.nfib));
  long long ctas = ((long long)cc.nslices * 32 + 255) / 256;
  const int grid = (int)(ctas < (1 << 20) ? (ctas > 0 ? ctas : 1) : (1 << 20));
  csf3_fiber_cells_kernel<<<grid, 256, 0, stream()>>>(cc.p1.as<int>(), cc.c1.as<int>(), cc.p2.as<int>(), cc.c2.as<int>(),
                                                     (unsigned)Kdim, (unsigned*)*cell);
  count_launch(1);
  return TACO_B200_OK;
}

// TTM = SpMM over the fibers: "row" f has the leaves [B3_pos[f], B3_pos[f+1]), gathers rows of C and is stored at row
// fiber_cell[f] of the (I*K) x R result.  The nnz-balanced slot kernel of spmm.cu keeps the reference's order (leaf
// order, separate multiply and add).  TACO_B200_TTM_VARIANT=1 selects the warp-per-fiber kernel this replaced.
template <typename T>
static int ttm_launch(CsfCall& cc, const T* C, T* A, size_t a_count, int R, int Kdim) {
  TB_CUDA(cudaMemsetAsync(A, 0, a_count * sizeof(T), stream()));
  static const int variant = getenv("TACO_B200_TTM_VARIANT") ? atoi(getenv("TACO_B200_TTM_VARIANT")) : 0;
  if (cc.nfib > 0 && R > 0 && variant == 0 && a_count / (size_t)R <= 0xFFFFFFFFull && cc.nnz <= INT32_MAX - 65536) {
    void* cell = nullptr;
    TB_TRY(fiber_cells(cc, Kdim, &cell));
    const int rc = spmm_mapped(sizeof(T) == 8 ? DType::F64 : DType::F32, cc.p3.as<int>(), cc.c3.as<int>(), cc.vals.as<T>(), C, A,
                               cc.nfib, cc.B.dim[2], R, cc.nnz, (const unsigned*)cell, "ttm_csf");
    scratch_free(cell);
    count_launch(1);
    TB_TRY(rc);
    TB_CUDA(cudaGetLastError());
    return TACO_B200_OK;
  }
  if (cc.nfib > 0 && R > 0) {
    long long ctas = ((long long)cc.nfib + CSF_WARPS - 1) / CSF_WARPS;
    int grid = (int)(ctas < (1 << 22) ? ctas : (1 << 22));
    ProfScope ps("ttm_csf");
    csf3_rows_kernel<T, 1><<<grid, CSF_WARPS * 32, 0, stream()>>>(cc.p1.as<int>(), cc.c1.as<int>(), cc.p2.as<int>(),
        cc.c2.as<int>(), cc.p3.as<int>(), cc.c3.as<int>(), cc.vals.as<T>(), C, (const T*)nullptr, A, R, Kdim, cc.B.dim[0]);
    count_launch(1);
  }
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

// TTV = SpMV over the fibers: "row" f has the leaves [B3_pos[f], B3_pos[f+1]) and is stored at A[fiber_cell[f]].  The
// nnz-balanced single-kernel SpMV (spmv.cu) keeps the reference's order (scalar accumulator, ascending leaf position).
// TACO_B200_TTV_VARIANT=1 selects the thread-per-fiber kernel this replaced.
template <typename T>
static int ttv_launch(CsfCall& cc, const T* c, T* A, size_t a_count, int Kdim) {
  TB_CUDA(cudaMemsetAsync(A, 0, a_count * sizeof(T), stream()));
  static const int variant = getenv("TACO_B200_TTV_VARIANT") ? atoi(getenv("TACO_B200_TTV_VARIANT")) : 0;
  if (cc.nfib > 0 && variant == 0 && a_count <= 0xFFFFFFFFull && cc.nnz <= INT32_MAX - 65536) {
    void* cell = nullptr;
    TB_TRY(fiber_cells(cc, Kdim, &cell));
    count_launch(1);
    const int rc = spmv_mapped(sizeof(T) == 8 ? DType::F64 : DType::F32, cc.p3.as<int>(), cc.c3.as<int>(), cc.vals.as<T>(), c, A,
                               cc.nfib, cc.nnz, (const unsigned*)cell, "ttv_csf");
    scratch_free(cell);
    TB_TRY(rc);
    TB_CUDA(cudaGetLastError());
    return TACO_B200_OK;
  }
  if (cc.nfib > 0) {
    long long ctas = ((long long)cc.nfib + 255) / 256;
    int grid = (int)(ctas < (1 << 22) ? ctas : (1 << 22));
    ProfScope ps("ttv_csf");
    csf3_ttv_kernel<T><<<grid, 256, 0, stream()>>>(cc.p1.as<int>(), cc.c1.as<int>(), cc.p2.as<int>(), cc.c2.as<int>(),
                                                   cc.p3.as<int>(), cc.c3.as<int>(), cc.vals.as<T>(), c, A, Kdim);
    count_launch(1);
  }
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_mttkrp_assemble(taco_tensor_t* A, taco_tensor_t*, taco_tensor_t*, taco_tensor_t*) {
  return dense_assemble(A, 2, "mttkrp");
}

int taco_b200_mttkrp_compute(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  TB_TRY(ensure_init());
  DenseView Av, Cv, Dv;
  TB_TRY(view_dense(A, 2, "A", &Av));
  TB_TRY(view_dense(C, 2, "C", &Cv));
  TB_TRY(view_dense(D, 2, "D", &Dv));
  CsfCall cc;
  TB_TRY(csf_prepare(B, &cc));
  if (Av.mode_order[0] != 0 || Cv.mode_order[0] != 0 || Dv.mode_order[0] != 0)
    return fail(TACO_B200_ERR_FORMAT, "mttkrp: A, C, D must be row-major {Dense,Dense}");
  const int R = Av.dim[1];
  if (Av.dim[0] != cc.B.dim[0] || Cv.dim[0] != cc.B.dim[1] || Dv.dim[0] != cc.B.dim[2] || Cv.dim[1] != R || Dv.dim[1] != R)
    return fail(TACO_B200_ERR_ARG, "mttkrp: dimension mismatch");
  if (Av.dt != cc.dt || Cv.dt != cc.dt || Dv.dt != cc.dt) return fail(TACO_B200_ERR_FORMAT, "mttkrp: mixed component types");
  size_t es = dsize(cc.dt);
  In cin, din; Out aout;
  TB_TRY(cin.acquire(Cv.vals, es * Cv.count()));
  TB_TRY(din.acquire(Dv.vals, es * Dv.count()));
  if (cc.host_described && classify(Av.vals) != Mem::Device && cc.nslices > 0 && cc.nnz > 0 && R > 0 && cc.nnz <= INT32_MAX - 65536 &&
      (4 + es) * (size_t)cc.nnz + es * Av.count() >= csf_pipeline_min_bytes()) {
    if (!Av.vals) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
    if (cc.dt == DType::F64) return mttkrp_compute_pipelined<double>(cc, cin.as<double>(), din.as<double>(), (double*)Av.vals, R);
    return mttkrp_compute_pipelined<float>(cc, cin.as<float>(), din.as<float>(), (float*)Av.vals, R);
  }
  TB_TRY(cc.upload());
  TB_TRY(aout.acquire(Av.vals, es * Av.count()));
  if (cc.dt == DType::F64) TB_TRY(mttkrp_launch<double>(cc, cin.as<double>(), din.as<double>(), aout.as<double>(), Av.count(), R));
  else TB_TRY(mttkrp_launch<float>(cc, cin.as<float>(), din.as<float>(), aout.as<float>(), Av.count(), R));
  TB_TRY(aout.commit());
  return finish_call();
}

int taco_b200_mttkrp_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  TB_TRY(taco_b200_mttkrp_assemble(A, B, C, D));
  return taco_b200_mttkrp_compute(A, B, C, D);
}

int taco_b200_ttv_assemble(taco_tensor_t* A, taco_tensor_t*, taco_tensor_t*) { return dense_assemble(A, 2, "ttv"); }

int taco_b200_ttv_compute(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* c) {
  TB_TRY(ensure_init());
  DenseView Av, cv;
  TB_TRY(view_dense(A, 2, "A", &Av));
  TB_TRY(view_dense(c, 1, "c", &cv));
  CsfCall cc;
  TB_TRY(csf_prepare(B, &cc));
  for (int l = 0; l < 3; l++)
    if (B->mode_ordering[l] != l) return fail(TACO_B200_ERR_FORMAT, "B must be stored in mode ordering 0,1,2 for this statement");
  TB_TRY(cc.upload());
  if (Av.mode_order[0] != 0) return fail(TACO_B200_ERR_FORMAT, "ttv: A must be row-major");
  if (Av.dim[0] != cc.B.dim[0] || Av.dim[1] != cc.B.dim[1] || cv.dim[0] != cc.B.dim[2])
    return fail(TACO_B200_ERR_ARG, "ttv: dimension mismatch");
  if (Av.dt != cc.dt || cv.dt != cc.dt) return fail(TACO_B200_ERR_FORMAT, "ttv: mixed component types");
  size_t es = dsize(cc.dt);
  In cin; Out aout;
  TB_TRY(cin.acquire(cv.vals, es * cv.count()));
  TB_TRY(aout.acquire(Av.vals, es * Av.count()));
  if (cc.dt == DType::F64) TB_TRY(ttv_launch<double>(cc, cin.as<double>(), aout.as<double>(), Av.count(), Av.dim[1]));
  else TB_TRY(ttv_launch<float>(cc, cin.as<float>(), aout.as<float>(), Av.count(), Av.dim[1]));
  TB_TRY(aout.commit());
  return finish_call();
}

int taco_b200_ttv_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* c) {
  TB_TRY(taco_b200_ttv_assemble(A, B, c));
  return taco_b200_ttv_compute(A, B, c);
}

int taco_b200_ttm_assemble(taco_tensor_t* A, taco_tensor_t*, taco_tensor_t*) { return dense_assemble(A, 3, "ttm"); }

int taco_b200_ttm_compute(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C) {
  TB_TRY(ensure_init());
  DenseView Av, Cv;
  TB_TRY(view_dense(A, 3, "A", &Av));
  TB_TRY(view_dense(C, 2, "C", &Cv));
  CsfCall cc;
  TB_TRY(csf_prepare(B, &cc));
  for (int l = 0; l < 3; l++)
    if (B->mode_ordering[l] != l) return fail(TACO_B200_ERR_FORMAT, "B must be stored in mode ordering 0,1,2 for this statement");
  TB_TRY(cc.upload());
  if (Av.mode_order[0] != 0 || Av.mode_order[1] != 1 || Cv.mode_order[0] != 0)
    return fail(TACO_B200_ERR_FORMAT, "ttm: A and C must be row-major");
  const int R = Av.dim[2];
  if (Av.dim[0] != cc.B.dim[0] || Av.dim[1] != cc.B.dim[1] || Cv.dim[0] != cc.B.dim[2] || Cv.dim[1] != R)
    return fail(TACO_B200_ERR_ARG, "ttm: dimension mismatch");
  if (Av.dt != cc.dt || Cv.dt != cc.dt) return fail(TACO_B200_ERR_FORMAT, "ttm: mixed component types");
  size_t es = dsize(cc.dt);
  In cin; Out aout;
  TB_TRY(cin.acquire(Cv.vals, es * Cv.count()));
  TB_TRY(aout.acquire(Av.vals, es * Av.count()));
  if (cc.dt == DType::F64) TB_TRY(ttm_launch<double>(cc, cin.as<double>(), aout.as<double>(), Av.count(), R, Av.dim[1]));
  else TB_TRY(ttm_launch<float>(cc, cin.as<float>(), aout.as<float>(), Av.count(), R, Av.dim[1]));
  TB_TRY(aout.commit());
  return finish_call();
}

int taco_b200_ttm_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C) {
  TB_TRY(taco_b200_ttm_assemble(A, B, C));
  return taco_b200_ttm_compute(A, B, C);
}

}  // extern "C"
