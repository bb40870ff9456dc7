// sparse_out.cu -- GPU assembly of sparse results: SpAdd  C(i,j) = A(i,j) + B(i,j)  and
//                                                   SpGEMM C(i,k) = A(i,j) * B(j,k),   all CSR.
//
// The reference has NO GPU path for sparse outputs: workspaces are disabled under CUDA
// (/root/reference/src/lower/lowerer_impl_imperative.cpp:2286-2290) and assembly is a host-serial append loop with
// realloc-by-copy on managed memory (src/codegen/codegen_cuda.cpp:1071-1135).  This file is the "src/lower gains GPU
// assembly" item of the north star: the two-phase Insert strategy the reference lowers for CPUs
// (lowerAssemble :2616-2779; CompressedModeFormat getSeqInitEdges/getSeqInsertEdge/getYieldPos/getFinalizeYieldPos,
// src/lower/mode_format_compressed.cpp:217-271) executed on the device:
//     symbolic   per-row size of the result pattern            (kernel, rows in parallel)
//     scan       pos = exclusive prefix sum of the sizes        (scan.cuh)
//     fill       crd written in ascending column order          (kernel)
//     numeric    values written at the positions pos/crd define (kernel; `compute`)
// Structure rules that make pos/crd BIT-EXACT with the reference: SpAdd = two-finger union keeping explicit zeros;
// SpGEMM = sorted set of reachable columns per row, entries that sum to zero are kept.
// SpGEMM values are accumulated in the reference's order (A-row order, then B-row order) -> bit-identical.
#include <climits>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "scan.cuh"

namespace tb {

int csr_nnz(const CsrView& A, int32_t vals_size_hint, int32_t* nnz);   // spmv.cu

// =========================================================================================================
// SpAdd
// =========================================================================================================
// One thread per row: both operand rows are short contiguous runs, so a warp touches one contiguous window of each
// crd/vals array (L1 absorbs the per-thread strides).
__global__ void __launch_bounds__(256)
spadd_count_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const int* __restrict__ Bpos,
                   const int* __restrict__ Bcrd, int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { counts[n] = 0; return; }
  int a = __ldg(Apos + i), ae = __ldg(Apos + i + 1), b = __ldg(Bpos + i), be = __ldg(Bpos + i + 1), c = 0;
  if (a < ae && b < be) {
    int ja = __ldg(Acrd + a), jb = __ldg(Bcrd + b);
    while (true) {
      c++;
      int j = min(ja, jb);
      if (ja == j) { if (++a == ae) { if (jb == j) ++b; break; } ja = __ldg(Acrd + a); }
      if (jb == j) { if (++b == be) break; jb = __ldg(Bcrd + b); }
    }
  }
  counts[i] = c + (ae - a) + (be - b);
}

template <typename T, bool CRD, bool VALS>
__global__ void __launch_bounds__(256)
spadd_fill_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const T* __restrict__ Av,
                  const int* __restrict__ Bpos, const int* __restrict__ Bcrd, const T* __restrict__ Bv,
                  const int* __restrict__ Cpos, int* __restrict__ Ccrd, T* __restrict__ Cv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = __ldg(Apos + i), ae = __ldg(Apos + i + 1), b = __ldg(Bpos + i), be = __ldg(Bpos + i + 1), p = __ldg(Cpos + i);
  while (a < ae && b < be) {
    int ja = __ldg(Acrd + a), jb = __ldg(Bcrd + b), j = min(ja, jb);
    if (CRD) Ccrd[p] = j;
    if (VALS) Cv[p] = (ja == j && jb == j) ? (__ldg(Av + a) + __ldg(Bv + b)) : (ja == j ? __ldg(Av + a) : __ldg(Bv + b));
    p++;
    a += (ja == j);
    b += (jb == j);
  }
  for (; a < ae; a++, p++) { if (CRD) Ccrd[p] = __ldg(Acrd + a); if (VALS) Cv[p] = __ldg(Av + a); }
  for (; b < be; b++, p++) { if (CRD) Ccrd[p] = __ldg(Bcrd + b); if (VALS) Cv[p] = __ldg(Bv + b); }
}

// Row-block kernels: a CTA owns `rb` consecutive rows.  Their segments of A and B (and of C) are contiguous, so they
// are moved between HBM and shared memory with fully coalesced accesses and the per-row two-finger merges run out of
// shared memory.  MODE: 0 = count (symbolic), 1 = crd only (assemble), 2 = vals only (compute), 3 = crd + vals
// (evaluate).  A row block whose segments exceed the staging capacity takes the direct path below.
// SPADD_CAP = staged entries per operand per CTA, SPADD_THREADS = staging threads (the first rb of them also merge one
// row each).  Small CTAs (512 entries, 64 threads, 24 KB for the fused fp64 fill) keep 9 CTAs per SM in different phases
// (load / merge / store), which is what hides the phase latencies.
template <typename T, int MODE, int SPADD_CAP, int SPADD_THREADS>
__global__ void __launch_bounds__(SPADD_THREADS)
spadd_block_kernel(int n, int rb, const int* __restrict__ Apos, const int* __restrict__ Acrd, const T* __restrict__ Av,
                   const int* __restrict__ Bpos, const int* __restrict__ Bcrd, const T* __restrict__ Bv,
                   const int* __restrict__ Cpos, int* __restrict__ Ccrd, T* __restrict__ Cv, int* __restrict__ counts) {
  constexpr bool CRD = (MODE & 1) != 0, VALS = (MODE & 2) != 0, COUNT = MODE == 0;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [A crd | B crd | C crd (if CRD)] ints, then [A vals | B vals | C vals] (if VALS)
  int* sAc = (int*)smem_raw;
  int* sBc = sAc + SPADD_CAP;
  int* sCc = sBc + SPADD_CAP;
  T* sAv = (T*)(sCc + (CRD ? 2 * SPADD_CAP : 0));
  T* sBv = sAv + SPADD_CAP;
  T* sCv = sBv + SPADD_CAP;
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * rb, r1 = min(r0 + rb, n);
  if (COUNT && blockIdx.x == gridDim.x - 1 && tid == 0) counts[n] = 0;
  const int a0 = __ldg(Apos + r0), a1 = __ldg(Apos + r1), b0 = __ldg(Bpos + r0), b1 = __ldg(Bpos + r1);
  const int r = r0 + tid;
  const bool mine = tid < rb && r < r1;
  if (a1 - a0 <= SPADD_CAP && b1 - b0 <= SPADD_CAP) {
#pragma unroll 4
    for (int q = tid; q < a1 - a0; q += SPADD_THREADS) { sAc[q] = tbd::ldg_stream_i32(Acrd + a0 + q); if (VALS) sAv[q] = __ldg(Av + a0 + q); }
#pragma unroll 4
    for (int q = tid; q < b1 - b0; q += SPADD_THREADS) { sBc[q] = tbd::ldg_stream_i32(Bcrd + b0 + q); if (VALS) sBv[q] = __ldg(Bv + b0 + q); }
    const int c0 = COUNT ? 0 : __ldg(Cpos + r0), c1 = COUNT ? 0 : __ldg(Cpos + r1);
    int a = 0, ae = 0, b = 0, be = 0, p = 0;
    if (mine) {
      a = __ldg(Apos + r) - a0; ae = __ldg(Apos + r + 1) - a0; b = __ldg(Bpos + r) - b0; be = __ldg(Bpos + r + 1) - b0;
      if (!COUNT) p = __ldg(Cpos + r) - c0;
    }
    __syncthreads();
    if (mine) {
      int cnt = 0;
      while (a < ae && b < be) {
        const int ja = sAc[a], jb = sBc[b], j = min(ja, jb);
        if (COUNT) cnt++;
        else {
          if (CRD) sCc[p] = j;
          if (VALS) sCv[p] = (ja == j && jb == j) ? (sAv[a] + sBv[b]) : (ja == j ? sAv[a] : sBv[b]);
          p++;
        }
        a += (ja == j);
        b += (jb == j);
      }
      if (COUNT) counts[r] = cnt + (ae - a) + (be - b);
      else {
        for (; a < ae; a++, p++) { if (CRD) sCc[p] = sAc[a]; if (VALS) sCv[p] = sAv[a]; }
        for (; b < be; b++, p++) { if (CRD) sCc[p] = sBc[b]; if (VALS) sCv[p] = sBv[b]; }
      }
    }
    if (!COUNT) {
      __syncthreads();
#pragma unroll 4
      for (int q = tid; q < c1 - c0; q += SPADD_THREADS) { if (CRD) Ccrd[c0 + q] = sCc[q]; if (VALS) Cv[c0 + q] = sCv[q]; }
    }
    return;
  }
  // ---- direct path: segments too long to stage ----------------------------------------------------------------------
  if (!mine) return;
  int a = __ldg(Apos + r), ae = __ldg(Apos + r + 1), b = __ldg(Bpos + r), be = __ldg(Bpos + r + 1);
  int p = COUNT ? 0 : __ldg(Cpos + r), cnt = 0;
  while (a < ae && b < be) {
    const int ja = __ldg(Acrd + a), jb = __ldg(Bcrd + b), j = min(ja, jb);
    if (COUNT) cnt++;
    else {
      if (CRD) Ccrd[p] = j;
      if (VALS) Cv[p] = (ja == j && jb == j) ? (__ldg(Av + a) + __ldg(Bv + b)) : (ja == j ? __ldg(Av + a) : __ldg(Bv + b));
      p++;
    }
    a += (ja == j);
    b += (jb == j);
  }
  if (COUNT) counts[r] = cnt + (ae - a) + (be - b);
  else {
    for (; a < ae; a++, p++) { if (CRD) Ccrd[p] = __ldg(Acrd + a); if (VALS) Cv[p] = __ldg(Av + a); }
    for (; b < be; b++, p++) { if (CRD) Ccrd[p] = __ldg(Bcrd + b); if (VALS) Cv[p] = __ldg(Bv + b); }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Element-parallel union kernel (default).  The two-finger merge above is serial per row; here every stored entry of
// the row block is one unit of work:
//   entry a of A(i,:)  ->  lb = #{entries of B(i,:) with column < j_a}   (binary search in the staged row of B)
//   entry b of B(i,:)  ->  ub = #{entries of A(i,:) with column <= j_b}; b is a DUPLICATE if A(i,:) holds j_b
// With A and B flattened over the row block, a's position in the union with duplicates is q_a + lb and the number of
// duplicates in front of any entry equals the number of MATCHED A entries in front of it -- an exclusive scan M over
// A's match flags -- so   final(a) = q_a + lb - M[q_a],   final(b) = q_b + ub - M[ub]   (duplicates dropped).
// The result is the reference's merge-lattice union (ascending columns, value a+b / a / b, explicit zeros kept),
// produced with coalesced loads, a CTA scan and coalesced stores.  MODE 0..3 as above; MODE 4 = one pass: tiles take
// tickets, publish their result count and obtain their offset in C by decoupled look-back, so pos, crd and vals are
// written by a single kernel that reads A and B exactly once (used by `evaluate` when the result arrays may be
// allocated at their upper bound nnzA + nnzB).
constexpr unsigned long long SPADD_AGG = 1ull << 32, SPADD_INCL = 2ull << 32;

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive offset of `tile` from the per-tile states (warp 0 calls this; every lane returns the same value)
__device__ __forceinline__ int spadd_look_back(const unsigned long long* state, int tile, int lane) {
  int excl = 0;
  for (int base = tile - 1; base >= 0; base -= 32) {
    const int p = base - lane;
    unsigned long long v = SPADD_INCL;               // virtual tiles in front of tile 0: inclusive prefix 0
    if (p >= 0) {
      do { v = ld_relaxed_u64(state + p); } while ((v >> 32) == 0);
    }
    const unsigned incl = __ballot_sync(0xffffffffu, (v >> 32) == 2);
    const int first = incl ? __ffs(incl) - 1 : 31;    // nearest predecessor that already knows its inclusive prefix
    int val = lane <= first ? (int)(unsigned)v : 0;
#pragma unroll
    for (int off = 16; off; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off);
    excl += val;
    if (incl) break;
  }
  return excl;
}

template <typename T, int MODE, int CAP, int THREADS>
__global__ void __launch_bounds__(THREADS)
spadd_union_kernel(int n, int rb, const int* __restrict__ Apos, const int* __restrict__ Acrd, const T* __restrict__ Av,
                   const int* __restrict__ Bpos, const int* __restrict__ Bcrd, const T* __restrict__ Bv,
                   int* __restrict__ Cpos, int* __restrict__ Ccrd, T* __restrict__ Cv, int* __restrict__ counts,
                   unsigned long long* __restrict__ tile_state, int* __restrict__ ticket) {
  constexpr int E = CAP / THREADS;
  static_assert(E == 4 && CAP == 4 * THREADS, "the CTA scan reads one int4 per thread");
  constexpr bool COUNT = MODE == 0, ONEPASS = MODE == 4;
  constexpr bool CRD = MODE == 1 || MODE == 3 || ONEPASS, VALS = MODE == 2 || MODE == 3 || ONEPASS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int* sF = (int*)smem_raw;                          // [CAP + 4] match flags of A -> exclusive scan (+ total)
  int* sAc = sF + CAP + 4;                           // [CAP]
  int* sBc = sAc + CAP;                              // [CAP]
  int* sAp = sBc + CAP;                              // [THREADS + 1] row starts of A relative to the block
  int* sBp = sAp + THREADS + 1;                      // [THREADS + 1]
  int* sCc = sBp + THREADS + 1 + ((2 * THREADS + 2) & 1);               // [2 CAP] if CRD (kept 8-byte aligned)
  unsigned short* sRa = (unsigned short*)(sCc + (CRD ? 2 * CAP : 0));  // [CAP] row of each A entry
  unsigned short* sRb = sRa + CAP;                                     // [CAP]
  T* sBv = (T*)(sRb + CAP);                          // [CAP] if VALS
  T* sCv = sBv + (VALS ? CAP : 0);                   // [2 CAP] if VALS
  __shared__ int s_tile, s_c0, s_warp[THREADS / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int tile = blockIdx.x;
  if (ONEPASS) {
    if (tid == 0) s_tile = atomicAdd(ticket, 1);
    __syncthreads();
    tile = s_tile;
  }
  const int r0 = tile * rb, r1 = min(r0 + rb, n), nr = r1 - r0;
  if (COUNT && tile == gridDim.x - 1 && tid == 0) counts[n] = 0;
  const int a0 = __ldg(Apos + r0), a1 = __ldg(Apos + r1), b0 = __ldg(Bpos + r0), b1 = __ldg(Bpos + r1);
  const int nA = a1 - a0, nB = b1 - b0;
  const bool staged = nA <= CAP && nB <= CAP;
  int my_as = 0, my_ae = 0, my_bs = 0, my_be = 0;
  if (tid < nr) {
    my_as = __ldg(Apos + r0 + tid) - a0; my_ae = __ldg(Apos + r0 + tid + 1) - a0;
    my_bs = __ldg(Bpos + r0 + tid) - b0; my_be = __ldg(Bpos + r0 + tid + 1) - b0;
  }
  if (staged) {
    // ---- stage: crd of both operands (and B's values) in shared memory, A's values in registers ------------------
    int ja[E], jb[E];
    T va[E], vb[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int q = tid + k * THREADS;
      ja[k] = 0; jb[k] = 0; va[k] = T(0); vb[k] = T(0);
      if (q < nA) { ja[k] = tbd::ldg_stream_i32(Acrd + a0 + q); if (VALS) va[k] = __ldg(Av + a0 + q); }
      if (q < nB) { jb[k] = tbd::ldg_stream_i32(Bcrd + b0 + q); if (VALS) vb[k] = __ldg(Bv + b0 + q); }
    }
    if (tid < nr) {
      sAp[tid] = my_as; sBp[tid] = my_bs;
      if (tid == nr - 1) { sAp[nr] = my_ae; sBp[nr] = my_be; }
      for (int q = my_as; q < my_ae; q++) sRa[q] = (unsigned short)tid;
      if (!COUNT) for (int q = my_bs; q < my_be; q++) sRb[q] = (unsigned short)tid;
    }
    if (COUNT) { if (tid < nr) sF[tid] = 0; }
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int q = tid + k * THREADS;
      if (q < nA) sAc[q] = ja[k];
      if (q < nB) { sBc[q] = jb[k]; if (VALS) sBv[q] = vb[k]; }
    }
    __syncthreads();
    // ---- A entries: rank among the entries of the same row of B, match flag ---------------------------------------
    int lbA[E];
    bool mA[E];
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int q = tid + k * THREADS;
      lbA[k] = 0; mA[k] = false;
      if (q < nA) {
        const int row = sRa[q];
        int lo = sBp[row], hi = sBp[row + 1];
        const int be = hi;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sBc[mid] < ja[k]) lo = mid + 1; else hi = mid; }
        lbA[k] = lo;
        mA[k] = lo < be && sBc[lo] == ja[k];
        if (COUNT && mA[k]) atomicAdd(sF + row, 1);
      }
      if (!COUNT) sF[q] = mA[k] ? 1 : 0;             // q < CAP always; entries beyond nA are zero
    }
    __syncthreads();
    if (COUNT) {
      if (tid < nr) counts[r0 + tid] = (my_ae - my_as) + (my_be - my_bs) - sF[tid];
      return;
    }
    // ---- exclusive scan of the match flags (one int4 per thread, warp shuffles, one partial per warp) -------------
    int4 f = ((int4*)sF)[tid];
    const int tsum = f.x + f.y + f.z + f.w;
    int incl = tsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) { const int v = s_warp[w]; if (w < wid) wbase += v; total += v; }
    int ex = wbase + incl - tsum;
    ((int4*)sF)[tid] = make_int4(ex, ex + f.x, ex + f.x + f.y, ex + f.x + f.y + f.z);
    if (tid == 0) sF[CAP] = total;
    const int nC = nA + nB - total;
    if (ONEPASS && tid == 0) st_relaxed_u64(tile_state + tile, (tile == 0 ? SPADD_INCL : SPADD_AGG) | (unsigned)nC);
    __syncthreads();
    // ---- scatter into the staged result ------------------------------------------------------------------------------
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int q = tid + k * THREADS;
      if (q < nA) {
        const int fin = q + lbA[k] - sF[q];
        if (CRD) sCc[fin] = ja[k];
        if (VALS) sCv[fin] = mA[k] ? va[k] + sBv[lbA[k]] : va[k];
      }
    }
#pragma unroll
    for (int k = 0; k < E; k++) {
      const int q = tid + k * THREADS;
      if (q < nB) {
        const int row = sRb[q];
        int lo = sAp[row], hi = sAp[row + 1];
        const int as = lo;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sAc[mid] <= jb[k]) lo = mid + 1; else hi = mid; }
        if (!(lo > as && sAc[lo - 1] == jb[k])) {
          const int fin = q + lo - sF[lo];
          if (CRD) sCc[fin] = jb[k];
          if (VALS) sCv[fin] = vb[k];
        }
      }
    }
    int c0;
    if (ONEPASS) {
      if (wid == 0) {
        const int excl = spadd_look_back(tile_state, tile, lane);
        if (lane == 0) {
          if (tile > 0) st_relaxed_u64(tile_state + tile, SPADD_INCL | (unsigned)(excl + nC));
          s_c0 = excl;
        }
      }
      __syncthreads();
      c0 = s_c0;
      if (tid < nr) {
        Cpos[r0 + tid] = c0 + my_as + my_bs - sF[my_as];
        if (r0 + tid == n - 1) Cpos[n] = c0 + nC;
      }
    } else {
      c0 = __ldg(Cpos + r0);
      __syncthreads();
    }
#pragma unroll 4
    for (int q = tid; q < nC; q += THREADS) { if (CRD) Ccrd[c0 + q] = sCc[q]; if (VALS) Cv[c0 + q] = sCv[q]; }
    return;
  }
  // ---- direct path: a row block too long to stage; one thread merges one row against global memory -------------------
  const bool mine = tid < nr;
  int a = a0 + my_as, ae = a0 + my_ae, b = b0 + my_bs, be = b0 + my_be;
  int p = 0;
  if (COUNT || ONEPASS) {
    int cnt = 0, x = a, y = b;
    if (mine) {
      while (x < ae && y < be) {
        const int jx = __ldg(Acrd + x), jy = __ldg(Bcrd + y), j = min(jx, jy);
        cnt++; x += (jx == j); y += (jy == j);
      }
      cnt += (ae - x) + (be - y);
    }
    if (COUNT) { if (mine) counts[r0 + tid] = cnt; return; }
    // one pass: CTA scan of the row sizes, publish, look back
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) { const int v = s_warp[w]; if (w < wid) wbase += v; total += v; }
    if (tid == 0) st_relaxed_u64(tile_state + tile, (tile == 0 ? SPADD_INCL : SPADD_AGG) | (unsigned)total);
    if (wid == 0) {
      const int excl = spadd_look_back(tile_state, tile, lane);
      if (lane == 0) {
        if (tile > 0) st_relaxed_u64(tile_state + tile, SPADD_INCL | (unsigned)(excl + total));
        s_c0 = excl;
      }
    }
    __syncthreads();
    p = s_c0 + wbase + incl - cnt;
    if (mine) {
      Cpos[r0 + tid] = p;
      if (r0 + tid == n - 1) Cpos[n] = p + cnt;
    }
  } else if (mine) {
    p = __ldg(Cpos + r0 + tid);
  }
  if (!mine) return;
  while (a < ae && b < be) {
    const int ja = __ldg(Acrd + a), jb = __ldg(Bcrd + b), j = min(ja, jb);
    if (CRD) Ccrd[p] = j;
    if (VALS) Cv[p] = (ja == j && jb == j) ? (__ldg(Av + a) + __ldg(Bv + b)) : (ja == j ? __ldg(Av + a) : __ldg(Bv + b));
    p++;
    a += (ja == j);
    b += (jb == j);
  }
  for (; a < ae; a++, p++) { if (CRD) Ccrd[p] = __ldg(Acrd + a); if (VALS) Cv[p] = __ldg(Av + a); }
  for (; b < be; b++, p++) { if (CRD) Ccrd[p] = __ldg(Bcrd + b); if (VALS) Cv[p] = __ldg(Bv + b); }
}

template <typename T, int MODE, int CAP, int THREADS>
static int spadd_union_launch_cfg(int n, int nnzA, int nnzB, const int* Apos, const int* Acrd, const T* Av, const int* Bpos,
                                  const int* Bcrd, const T* Bv, int* Cpos, int* Ccrd, T* Cv, int* counts) {
  constexpr bool ONEPASS = MODE == 4;
  constexpr bool CRD = MODE == 1 || MODE == 3 || ONEPASS, VALS = MODE == 2 || MODE == 3 || ONEPASS;
  const size_t smem = sizeof(int) * ((CAP + 4) + 2 * CAP + 2 * (THREADS + 1) + ((2 * THREADS + 2) & 1) + (CRD ? 2 * CAP : 0)) +
                      sizeof(unsigned short) * 2 * CAP + (VALS ? sizeof(T) * 3 * CAP : 0);
  static bool configured = false;
  if (!configured) {
    TB_CUDA(cudaFuncSetAttribute(spadd_union_kernel<T, MODE, CAP, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // rows per CTA: as many as keep an average segment at ~75% of the staging capacity (at most one per thread)
  const double avg = (double)(nnzA > nnzB ? nnzA : nnzB) / (n > 0 ? n : 1);
  int rb = avg > 0 ? (int)(0.75 * CAP / avg) : THREADS;
  rb = rb > THREADS ? THREADS : (rb < 4 ? 4 : rb);
  const int grid = (n + rb - 1) / rb;
  unsigned long long* state = nullptr;
  if (ONEPASS) {
    TB_TRY(scratch_alloc((void**)&state, sizeof(unsigned long long) * ((size_t)grid + 1)));
    TB_CUDA(cudaMemsetAsync(state, 0, sizeof(unsigned long long) * ((size_t)grid + 1), stream()));
  }
  spadd_union_kernel<T, MODE, CAP, THREADS><<<grid, THREADS, smem, stream()>>>(n, rb, Apos, Acrd, Av, Bpos, Bcrd, Bv, Cpos, Ccrd,
                                                                               Cv, counts, state, (int*)(state + grid));
  count_launch(1);
  if (ONEPASS) scratch_free(state);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

template <typename T, int MODE, int CAP, int THREADS>
static int spadd_block_launch_cfg(int n, int nnzA, int nnzB, const int* Apos, const int* Acrd, const T* Av, const int* Bpos,
                                  const int* Bcrd, const T* Bv, const int* Cpos, int* Ccrd, T* Cv, int* counts) {
  constexpr bool CRD = (MODE & 1) != 0, VALS = (MODE & 2) != 0;
  const size_t smem = sizeof(int) * CAP * (2 + (CRD ? 2 : 0)) + (VALS ? sizeof(T) * CAP * 4 : 0);
  static bool configured = false;
  if (!configured) {
    TB_CUDA(cudaFuncSetAttribute(spadd_block_kernel<T, MODE, CAP, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // rows per CTA: as many as keep an average segment at ~70% of the staging capacity (at most one per thread)
  const double avg = (double)(nnzA > nnzB ? nnzA : nnzB) / (n > 0 ? n : 1);
  int rb = avg > 0 ? (int)(0.7 * CAP / avg) : THREADS;
  rb = rb > THREADS ? THREADS : (rb < 4 ? 4 : rb);
  const int grid = (n + rb - 1) / rb;
  spadd_block_kernel<T, MODE, CAP, THREADS><<<grid, THREADS, smem, stream()>>>(n, rb, Apos, Acrd, Av, Bpos, Bcrd, Bv, Cpos, Ccrd,
                                                                               Cv, counts);
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

template <typename T, int MODE>
static int spadd_block_launch(int n, int nnzA, int nnzB, const int* Apos, const int* Acrd, const T* Av, const int* Bpos,
                              const int* Bcrd, const T* Bv, const int* Cpos, int* Ccrd, T* Cv, int* counts) {
  static const int variant = getenv("TACO_B200_SPADD_VARIANT") ? atoi(getenv("TACO_B200_SPADD_VARIANT")) : 0;
#define TB_SPADD_ARGS n, nnzA, nnzB, Apos, Acrd, Av, Bpos, Bcrd, Bv, Cpos, Ccrd, Cv, counts
#define TB_SPADD_UARGS n, nnzA, nnzB, Apos, Acrd, Av, Bpos, Bcrd, Bv, (int*)Cpos, Ccrd, Cv, counts
  switch (variant) {
    case 1: return spadd_block_launch_cfg<T, MODE, 2048, 256>(TB_SPADD_ARGS);
    case 2: return spadd_block_launch_cfg<T, MODE, 1024, 128>(TB_SPADD_ARGS);
    case 3: return spadd_block_launch_cfg<T, MODE, 256, 32>(TB_SPADD_ARGS);
    case 4: return spadd_block_launch_cfg<T, MODE, 512, 128>(TB_SPADD_ARGS);
    case 5: return spadd_block_launch_cfg<T, MODE, 512, 64>(TB_SPADD_ARGS);
    case 6: return spadd_union_launch_cfg<T, MODE, 1024, 256>(TB_SPADD_UARGS);
    case 7: return spadd_union_launch_cfg<T, MODE, 256, 64>(TB_SPADD_UARGS);
    default: return spadd_union_launch_cfg<T, MODE, 512, 128>(TB_SPADD_UARGS);
  }
#undef TB_SPADD_ARGS
}

// one-pass union (MODE 4): Cpos is an output
template <typename T>
static int spadd_onepass_launch(int n, int nnzA, int nnzB, const int* Apos, const int* Acrd, const T* Av, const int* Bpos,
                                const int* Bcrd, const T* Bv, int* Cpos, int* Ccrd, T* Cv) {
  static const int variant = getenv("TACO_B200_SPADD_VARIANT") ? atoi(getenv("TACO_B200_SPADD_VARIANT")) : 0;
  int* counts = nullptr;
  switch (variant) {
    case 6: return spadd_union_launch_cfg<T, 4, 1024, 256>(TB_SPADD_UARGS);
    case 7: return spadd_union_launch_cfg<T, 4, 256, 64>(TB_SPADD_UARGS);
    default: return spadd_union_launch_cfg<T, 4, 512, 128>(TB_SPADD_UARGS);
  }
}
#undef TB_SPADD_UARGS

// =========================================================================================================
// SpGEMM
// =========================================================================================================
constexpr int SG_WARP_CAP = 256;     // products per row handled by a warp team in shared memory
constexpr int SG_CTA_CAP = 8192;     // products per row handled by a CTA team in shared memory
constexpr int SG_BINS = 5;           // 0: <=64, 1: <=128, 2: <=256 products (warp, registers), 3: <=8192 (CTA), 4: bitmap

// upper bound of the row pattern = number of products; rows are binned by it
__global__ void __launch_bounds__(256)
spgemm_bound_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const int* __restrict__ Bpos,
                    int* __restrict__ bin_count, int* __restrict__ bin_rows, int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long ub = 0;
  for (int p = __ldg(Apos + i); p < __ldg(Apos + i + 1); p++) {
    int j = __ldg(Acrd + p);
    ub += __ldg(Bpos + j + 1) - __ldg(Bpos + j);
  }
  if (ub == 0) { if (counts) counts[i] = 0; return; }
  int bin = ub <= 64 ? 0 : (ub <= 128 ? 1 : (ub <= 256 ? 2 : (ub <= SG_CTA_CAP ? 3 : 4)));
  // warp-aggregated append: one atomic per (warp, bin)
  for (int b = 0; b < SG_BINS; b++) {
    const unsigned m = __ballot_sync(__activemask(), bin == b);
    if (bin == b) {
      const int leader = __ffs(m) - 1, lane = threadIdx.x & 31;
      int base = 0;
      if (lane == leader) base = atomicAdd(bin_count + b, __popc(m));
      base = __shfl_sync(m, base, leader);
      bin_rows[(size_t)b * n + base + __popc(m & ((1u << lane) - 1))] = i;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Rows with at most EPL*32 products: one warp per row, everything in registers / per-warp shared memory.
// ---------------------------------------------------------------------------------------------------------
// Expansion shared by the count and the fill kernel: visit every product (t, k, a*b) of row i in generation order
// t = 0,1,.. (A entries ascending, then the B row in order -- the order of the reference's workspace loop, Appendix
// A.5).  G lanes share one A entry, so short B rows still fill the warp.
template <typename T, bool VALS, bool UNIFORM, typename F>
__device__ __forceinline__ int spgemm_expand(int i, int G, const int* __restrict__ Apos, const int* __restrict__ Acrd,
                                             const T* __restrict__ Av, const int* __restrict__ Bpos,
                                             const int* __restrict__ Bcrd, const T* __restrict__ Bv, F&& visit) {
  const int lane = threadIdx.x & 31;
  const int grp = lane / G, gl = lane % G, NG = 32 / G;
  const int a0 = __ldg(Apos + i), a1 = __ldg(Apos + i + 1);
  int total = 0;
  for (int ab = a0; ab < a1; ab += 32) {
    int bs = 0, len = 0;
    T av = T(0);
    if (ab + lane < a1) {
      const int j = __ldg(Acrd + ab + lane);
      if (VALS) av = __ldg(Av + ab + lane);
      bs = __ldg(Bpos + j);
      len = __ldg(Bpos + j + 1) - bs;
    }
    int incl = len;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += v;
    }
    const int off = total + incl - len;
    const int cnt = min(32, a1 - ab);
    for (int q0 = 0; q0 < cnt; q0 += NG) {
      const int q = min(q0 + grp, 31);
      const int qbs = __shfl_sync(0xffffffffu, bs, q);
      int qlen = __shfl_sync(0xffffffffu, len, q);
      if (q0 + grp >= cnt) qlen = 0;
      const int qoff = __shfl_sync(0xffffffffu, off, q);
      const T qa = __shfl_sync(0xffffffffu, av, q);
      if (UNIFORM) {
        // warp-uniform trip count (visit() uses warp barriers): lanes past the end of their B row pass valid=false
        const int iters = (__reduce_max_sync(0xffffffffu, qlen) + G - 1) / G;
        for (int it = 0; it < iters; it++) {
          const int l = gl + it * G;
          const bool valid = l < qlen;
          const int k = valid ? __ldg(Bcrd + qbs + l) : -1;
          visit(qoff + l, k, T(0), valid);
        }
      } else {
        for (int l = gl; l < qlen; l += G) {
          const int k = __ldg(Bcrd + qbs + l);
          T prod = T(0);
          if (VALS) prod = qa * __ldg(Bv + qbs + l);
          visit(qoff + l, k, prod, true);
        }
      }
    }
    total += __shfl_sync(0xffffffffu, incl, 31);
  }
  return total;
}

// symbolic count: distinct columns among the products of a row, with a per-warp open-addressing hash set
template <int EPL>
__global__ void __launch_bounds__(256)
spgemm_count_warp_kernel(const int* __restrict__ rows_list, int nrows_bin, int G, const int* __restrict__ Apos,
                         const int* __restrict__ Acrd, const int* __restrict__ Bpos, const int* __restrict__ Bcrd,
                         int* __restrict__ counts) {
  constexpr int SLOTS = EPL * 64;                 // 2x the product capacity
  __shared__ int table_all[8][SLOTS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ridx = blockIdx.x * 8 + wid;
  if (ridx >= nrows_bin) return;
  int* table = table_all[wid];
#pragma unroll
  for (int t = 0; t < SLOTS / 32; t++) table[t * 32 + lane] = -1;
  __syncwarp();
  const int i = __ldg(rows_list + ridx);
  int mine = 0;
  // (A barrier-based insert without shared-memory atomics was measured slower: 0.99 vs 0.70 ms at C5.)
  spgemm_expand<float, false, false>(i, G, Apos, Acrd, (const float*)nullptr, Bpos, Bcrd, (const float*)nullptr,
                                     [&](int, int k, float, bool) {
                                       unsigned h = ((unsigned)k * 2654435761u) >> 7;
                                       while (true) {
                                         h &= SLOTS - 1;
                                         const int old = atomicCAS(table + h, -1, k);
                                         if (old == -1) { mine++; break; }
                                         if (old == k) break;
                                         h++;
                                       }
                                     });
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, off);
  if (lane == 0) counts[i] = mine;
}

// fill: expand (column<<8 | t) keys, bitonic-sort them in registers (EPL per lane, blocked layout: distances < EPL are
// in-lane, the rest are shuffles), then every run of equal columns is summed IN GENERATION ORDER by the lane that holds
// its head -- the first product is stored, later ones added, like the reference's `w[k] = a*b` / `w[k] += a*b`.
template <typename T, typename KEY, int EPL, bool CRD, bool VALS>
__global__ void __launch_bounds__(256)
spgemm_fill_warp_kernel(const int* __restrict__ rows_list, int nrows_bin, int G, const int* __restrict__ Apos,
                        const int* __restrict__ Acrd, const T* __restrict__ Av, const int* __restrict__ Bpos,
                        const int* __restrict__ Bcrd, const T* __restrict__ Bv, const int* __restrict__ Cpos,
                        int* __restrict__ Ccrd, T* __restrict__ Cv) {
  constexpr int N = EPL * 32;
  __shared__ __align__(16) KEY skey_all[8][N];
  __shared__ __align__(16) T sval_all[8][VALS ? N : 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int ridx = blockIdx.x * 8 + wid;
  if (ridx >= nrows_bin) return;
  KEY* skey = skey_all[wid];
  T* sval = sval_all[wid];
  const int i = __ldg(rows_list + ridx);
  const int total = spgemm_expand<T, VALS, false>(i, G, Apos, Acrd, Av, Bpos, Bcrd, Bv, [&](int t, int k, T prod, bool valid) {
    if (valid) {
      skey[t] = ((KEY)(unsigned)k << 8) | (KEY)t;
      if (VALS) sval[t] = prod;
    }
  });
  for (int t = total + lane; t < N; t += 32) skey[t] = ~(KEY)0;
  __syncwarp();
  KEY key[EPL];
#pragma unroll
  for (int e = 0; e < EPL; e++) key[e] = skey[lane * EPL + e];
  // bitonic sort, ascending over x = lane*EPL + e
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= EPL) {
        const int lj = j / EPL;
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          const int x = lane * EPL + e;
          const KEY other = __shfl_xor_sync(0xffffffffu, key[e], lj);
          const bool keep_min = ((x & k) == 0) == ((lane & lj) == 0);
          const KEY lo_ = key[e] < other ? key[e] : other, hi_ = key[e] < other ? other : key[e];
          key[e] = keep_min ? lo_ : hi_;
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPL; e++) {
          if ((e & j) == 0) {
            const int x = lane * EPL + e;
            const bool asc = (x & k) == 0;
            const KEY a = key[e], b = key[e | j];
            const bool sw = asc ? (a > b) : (a < b);
            key[e] = sw ? b : a;
            key[e | j] = sw ? a : b;
          }
        }
      }
    }
  }
  __syncwarp();
#pragma unroll
  for (int e = 0; e < EPL; e++) skey[lane * EPL + e] = key[e];
  __syncwarp();
  // run heads (striped positions p = e*32 + lane so that output writes are coalesced)
  const int c0 = __ldg(Cpos + i);
  int before = 0;
#pragma unroll
  for (int e = 0; e < EPL; e++) {
    const int p = e * 32 + lane;
    const KEY kp = skey[p];
    const bool head = p < total && (p == 0 || (skey[p - 1] >> 8) != (kp >> 8));
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (head) {
      const int rank = before + __popc(m & ((1u << lane) - 1));
      if (CRD) Ccrd[c0 + rank] = (int)(kp >> 8);
      if (VALS) {
        T acc = sval[(int)(kp & 255)];
        for (int pp = p + 1; pp < total && (skey[pp] >> 8) == (kp >> 8); pp++) acc = acc + sval[(int)(skey[pp] & 255)];
        Cv[c0 + rank] = acc;
      }
    }
    before += __popc(m);
  }
}

// Team-cooperative symbolic phase: expand the columns of all products of a row into shared memory, bitonic-sort them,
// count (WRITE=false) or emit (WRITE=true) the distinct ones in ascending order.
template <int THREADS, int CAP, bool WRITE>
__global__ void __launch_bounds__(256)
spgemm_symbolic_kernel(const int* __restrict__ rows_list, int nrows_bin, const int* __restrict__ Apos,
                       const int* __restrict__ Acrd, const int* __restrict__ Bpos, const int* __restrict__ Bcrd,
                       int* __restrict__ counts, const int* __restrict__ Cpos, int* __restrict__ Ccrd) {
  constexpr int TEAMS = 256 / THREADS;
  constexpr int CHUNK = CAP / THREADS;              // keys per thread in the compaction phase
  __shared__ int keys_all[TEAMS][CAP];
  __shared__ int aux_all[TEAMS][THREADS + 1];
  const int team = threadIdx.x / THREADS, tid = threadIdx.x % THREADS;
  const int ridx = blockIdx.x * TEAMS + team;
  auto team_sync = [&]() { if (THREADS == 32) __syncwarp(); else __syncthreads(); };
  // (a CTA team never exits early: TEAMS == 1 and the grid is exact; warp teams may return as a whole warp)
  if (ridx >= nrows_bin) return;
  int* keys = keys_all[team];
  int* aux = aux_all[team];
  const int i = __ldg(rows_list + ridx);
  const int a0 = __ldg(Apos + i), a1 = __ldg(Apos + i + 1);
  // ---- expand: THREADS A-entries at a time; a team-wide scan of the B-row lengths gives the write offsets -------
  int total = 0;
  for (int ab = a0; ab < a1; ab += THREADS) {
    int bs = 0, len = 0;
    if (ab + tid < a1) {
      int j = __ldg(Acrd + ab + tid);
      bs = __ldg(Bpos + j);
      len = __ldg(Bpos + j + 1) - bs;
    }
    aux[tid + 1] = len;
    team_sync();
    if (tid == 0) {
      aux[0] = 0;
      for (int t = 1; t <= THREADS; t++) aux[t] += aux[t - 1];
    }
    team_sync();
    const int off = total + aux[tid];
    // each thread copies its own B row: rows are short at the target shapes; long rows still coalesce per thread
    for (int t = 0; t < len; t++) keys[off + t] = __ldg(Bcrd + bs + t);
    total += aux[THREADS];
    team_sync();
  }
  int m = 32;
  while (m < total) m <<= 1;
  for (int t = total + tid; t < m; t += THREADS) keys[t] = INT_MAX;
  team_sync();
  // ---- bitonic sort of m keys -----------------------------------------------------------------------------
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (m >> 1); t += THREADS) {
        int lo = 2 * j * (t / j) + (t % j), hi = lo + j;
        int x = keys[lo], y = keys[hi];
        bool up = (lo & k) == 0;
        if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
      }
      team_sync();
    }
  }
  // ---- distinct keys: contiguous chunk per thread, team scan of the chunk counts -------------------------------
  const int c0 = tid * CHUNK;
  int cnt = 0;
  for (int t = c0; t < c0 + CHUNK && t < total; t++) cnt += (t == 0 || keys[t] != keys[t - 1]);
  if (!WRITE) {
    aux[tid] = cnt;
    team_sync();
    if (tid == 0) {
      int s = 0;
      for (int t = 0; t < THREADS; t++) s += aux[t];
      counts[i] = s;
    }
  } else {
    aux[tid + 1] = cnt;
    team_sync();
    if (tid == 0) {
      aux[0] = 0;
      for (int t = 1; t <= THREADS; t++) aux[t] += aux[t - 1];
    }
    team_sync();
    int* out = Ccrd + __ldg(Cpos + i) + aux[tid];
    for (int t = c0; t < c0 + CHUNK && t < total; t++)
      if (t == 0 || keys[t] != keys[t - 1]) *out++ = keys[t];
  }
}

// Rows with more than SG_CTA_CAP products: a bitmap over the column space in global scratch (one per CTA) gives the
// sorted distinct columns directly -- the device form of the reference's w_already_set / w_index_list + qsort.
template <bool WRITE>
__global__ void __launch_bounds__(256)
spgemm_symbolic_bitmap_kernel(const int* __restrict__ rows_list, int nrows_bin, int ncols, const int* __restrict__ Apos,
                              const int* __restrict__ Acrd, const int* __restrict__ Bpos, const int* __restrict__ Bcrd,
                              unsigned* __restrict__ bitmaps, int* __restrict__ counts, const int* __restrict__ Cpos,
                              int* __restrict__ Ccrd) {
  __shared__ int wsum[8];
  __shared__ int running;
  const int words = (ncols + 31) / 32;
  unsigned* bm = bitmaps + (size_t)blockIdx.x * words;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int ridx = blockIdx.x; ridx < nrows_bin; ridx += gridDim.x) {
    const int i = __ldg(rows_list + ridx);
    for (int t = threadIdx.x; t < words; t += 256) bm[t] = 0u;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int pa = __ldg(Apos + i) + wid; pa < __ldg(Apos + i + 1); pa += 8) {   // one warp per A entry
      const int j = __ldg(Acrd + pa);
      for (int pb = __ldg(Bpos + j) + lane; pb < __ldg(Bpos + j + 1); pb += 32) {
        int k = __ldg(Bcrd + pb);
        atomicOr(bm + (k >> 5), 1u << (k & 31));
      }
    }
    __syncthreads();
    for (int wb = 0; wb < words; wb += 256) {
      const int t = wb + threadIdx.x;
      const unsigned bits = t < words ? bm[t] : 0u;
      const int c = __popc(bits);
      int incl = c;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (lane == 31) wsum[wid] = incl;
      __syncthreads();
      int before = running;
      for (int q = 0; q < wid; q++) before += wsum[q];
      if (WRITE) {
        int* out = Ccrd + __ldg(Cpos + i) + before + incl - c;
        unsigned b = bits;
        while (b) { int bit = __ffs(b) - 1; *out++ = t * 32 + bit; b &= b - 1; }
      }
      __syncthreads();
      if (threadIdx.x == 255) running = before + incl;
      __syncthreads();
    }
    if (!WRITE && threadIdx.x == 0) counts[i] = running;
    __syncthreads();
  }
}

// Numeric phase.  A team of TEAM lanes owns a row.  A-entries are taken one after the other (the reference's outer
// order); the lanes of the team take the entries of that B row (distinct columns => no write conflicts), locate the
// column in the row's sorted crd by binary search and accumulate into C_vals -- first touch stores a*b, later touches
// add, exactly as the reference's workspace does (w[k] = a*b | w[k] = w[k] + a*b).
template <typename T, int TEAM>
__global__ void __launch_bounds__(256)
spgemm_numeric_kernel(const int* __restrict__ rows_list, int n, const int* __restrict__ Apos, const int* __restrict__ Acrd,
                      const T* __restrict__ Av, const int* __restrict__ Bpos, const int* __restrict__ Bcrd,
                      const T* __restrict__ Bv, const int* __restrict__ Cpos, const int* __restrict__ Ccrd, T* Cv) {
  const int gt = blockIdx.x * 256 + threadIdx.x;
  const int ridx = gt / TEAM, tl = gt % TEAM;
  const unsigned mask = TEAM == 32 ? 0xffffffffu : (((1u << TEAM) - 1u) << ((threadIdx.x & 31) / TEAM * TEAM));
  if (ridx >= n) return;
  const int i = rows_list ? __ldg(rows_list + ridx) : ridx;
  const int c0 = __ldg(Cpos + i), c1 = __ldg(Cpos + i + 1);
  for (int t = c0 + tl; t < c1; t += TEAM) Cv[t] = T(0);
  __syncwarp(mask);
  const int a0 = __ldg(Apos + i), a1 = __ldg(Apos + i + 1);
  for (int ab = a0; ab < a1; ab += TEAM) {
    int my_bs = 0, my_be = 0;
    T my_a = T(0);
    if (ab + tl < a1) {
      int j = __ldg(Acrd + ab + tl);
      my_a = __ldg(Av + ab + tl);
      my_bs = __ldg(Bpos + j);
      my_be = __ldg(Bpos + j + 1);
    }
    const int cnt = min(TEAM, a1 - ab);
    for (int q = 0; q < cnt; q++) {
      const int src = (threadIdx.x & 31) / TEAM * TEAM + q;
      const int bs = __shfl_sync(mask, my_bs, src), be = __shfl_sync(mask, my_be, src);
      const T a = __shfl_sync(mask, my_a, src);
      for (int pb = bs + tl; pb < be; pb += TEAM) {
        const int k = __ldg(Bcrd + pb);
        int lo = c0, hi = c1 - 1;                  // k is guaranteed present in Ccrd[c0..c1)
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (__ldg(Ccrd + mid) < k) lo = mid + 1; else hi = mid;
        }
        Cv[lo] = Cv[lo] + a * __ldg(Bv + pb);
      }
      __syncwarp(mask);
    }
  }
}

// =========================================================================================================
// host side
// =========================================================================================================
struct Csr3 { CsrView C, A, B; int32_t nnzA, nnzB; In apos, acrd, bpos, bcrd; };

static int csr3_prepare(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, bool product, Csr3* s) {
  TB_TRY(ensure_init());
  TB_TRY(view_csr(C, "C", &s->C));
  TB_TRY(view_csr(A, "A", &s->A));
  TB_TRY(view_csr(B, "B", &s->B));
  if (product) {
    if (s->A.cols != s->B.rows || s->C.rows != s->A.rows || s->C.cols != s->B.cols)
      return fail(TACO_B200_ERR_ARG, "spgemm: dimension mismatch");
  } else if (s->A.rows != s->B.rows || s->A.cols != s->B.cols || s->C.rows != s->A.rows || s->C.cols != s->A.cols) {
    return fail(TACO_B200_ERR_ARG, "spadd: dimension mismatch");
  }
  if (s->A.dt != s->B.dt || s->C.dt != s->A.dt) return fail(TACO_B200_ERR_FORMAT, "mixed component types");
  TB_TRY(csr_nnz(s->A, A->vals_size, &s->nnzA));
  TB_TRY(csr_nnz(s->B, B->vals_size, &s->nnzB));
  TB_TRY(s->apos.acquire(s->A.pos, sizeof(int32_t) * ((size_t)s->A.rows + 1)));
  TB_TRY(s->acrd.acquire(s->A.crd ? (void*)s->A.crd : (void*)s->A.pos, sizeof(int32_t) * (size_t)s->nnzA));
  TB_TRY(s->bpos.acquire(s->B.pos, sizeof(int32_t) * ((size_t)s->B.rows + 1)));
  TB_TRY(s->bcrd.acquire(s->B.crd ? (void*)s->B.crd : (void*)s->B.pos, sizeof(int32_t) * (size_t)s->nnzB));
  return TACO_B200_OK;
}

// hand freshly built device pos / crd (and, after a fused evaluate, vals) to the caller in the configured result space
static int publish_structure(taco_tensor_t* C, int n, int* dpos, int* dcrd, int32_t nnzC, size_t esize, void* dvals = nullptr) {
  if (result_space() == TACO_B200_SPACE_DEVICE) {
    void* vals = dvals;
    if (!vals) TB_TRY(device_result_alloc(&vals, esize * (size_t)(nnzC > 0 ? nnzC : 1)));
    C->indices[1][0] = (uint8_t*)dpos;
    C->indices[1][1] = (uint8_t*)dcrd;
    C->vals = (uint8_t*)vals;
  } else {
    int32_t* hpos = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* hcrd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnzC > 0 ? nnzC : 1));
    void* vals = malloc(esize * (size_t)(nnzC > 0 ? nnzC : 1));
    int rc = (!hpos || !hcrd || !vals) ? fail(TACO_B200_ERR_ALLOC, "cannot allocate host result arrays") : TACO_B200_OK;
    if (rc == TACO_B200_OK) rc = d2h_fresh(hpos, dpos, sizeof(int32_t) * ((size_t)n + 1));
    if (rc == TACO_B200_OK && nnzC) rc = d2h_fresh(hcrd, dcrd, sizeof(int32_t) * (size_t)nnzC);
    if (rc == TACO_B200_OK && nnzC && dvals) rc = d2h_fresh(vals, dvals, esize * (size_t)nnzC);
    if (rc == TACO_B200_OK && cudaStreamSynchronize(stream()) != cudaSuccess) rc = fail(TACO_B200_ERR_CUDA, "stream synchronisation failed");
    if (rc != TACO_B200_OK) { free(hpos); free(hcrd); free(vals); return rc; }
    device_result_free(dpos);
    device_result_free(dcrd);
    device_result_free(dvals);
    C->indices[1][0] = (uint8_t*)hpos;
    C->indices[1][1] = (uint8_t*)hcrd;
    C->vals = (uint8_t*)vals;
  }
  C->vals_size = nnzC;
  return TACO_B200_OK;
}

// symbolic count -> scan -> (host learns nnz) -> fill.  with_vals: the fill also writes the values (evaluate), so the
// operands are merged twice instead of three times.
template <typename T>
static int spadd_assemble_impl(taco_tensor_t* C, Csr3& s, const T* av, const T* bv) {
  const int n = s.A.rows;
  const bool with_vals = av != nullptr;
  int* dpos = nullptr;
  TB_TRY(device_result_alloc((void**)&dpos, sizeof(int) * ((size_t)n + 1)));
  // Default: two-phase (count -> scan -> fused crd+vals fill), measured 0.247 ms per C5a step.  TACO_B200_SPADD_ONEPASS=1
  // selects the single-kernel variant (tickets + decoupled look-back, arrays allocated at the upper bound nnzA + nnzB):
  // it reads A and B once, but the look-back spin costs more than the second read saves (0.285 ms, profiles/r01_variants.md).
  static const bool onepass_on = getenv("TACO_B200_SPADD_ONEPASS") && atoi(getenv("TACO_B200_SPADD_ONEPASS")) == 1 &&
                                 !(getenv("TACO_B200_SPADD_VARIANT") && atoi(getenv("TACO_B200_SPADD_VARIANT")) >= 1 &&
                                   atoi(getenv("TACO_B200_SPADD_VARIANT")) <= 5);
  const size_t bound = (size_t)s.nnzA + (size_t)s.nnzB;
  if (with_vals && onepass_on && n > 0 && bound > 0 && bound <= (size_t)INT32_MAX && bound * (4 + sizeof(T)) <= ((size_t)8 << 30)) {
    int* dcrd = nullptr;
    void* dvals = nullptr;
    TB_TRY(device_result_alloc((void**)&dcrd, sizeof(int) * bound));
    TB_TRY(device_result_alloc(&dvals, sizeof(T) * bound));
    {
      ProfScope ps("spadd_numeric");
      TB_TRY(spadd_onepass_launch<T>(n, s.nnzA, s.nnzB, s.apos.as<int>(), s.acrd.as<int>(), av, s.bpos.as<int>(),
                                     s.bcrd.as<int>(), bv, dpos, dcrd, (T*)dvals));
    }
    int32_t nnzC = 0;
    TB_TRY(read_back(&nnzC, dpos + n, sizeof(int32_t)));
    return publish_structure(C, n, dpos, dcrd, nnzC, sizeof(T), dvals);
  }
  if (n > 0) {
    ProfScope ps("spadd_symbolic");
    TB_TRY((spadd_block_launch<T, 0>(n, s.nnzA, s.nnzB, s.apos.as<int>(), s.acrd.as<int>(), nullptr, s.bpos.as<int>(),
                                     s.bcrd.as<int>(), nullptr, nullptr, nullptr, nullptr, dpos)));
  } else {
    TB_CUDA(cudaMemsetAsync(dpos, 0, sizeof(int), stream()));
  }
  int32_t nnzC = 0;             // 64-bit total of the row counts first: refuses results beyond int32 instead of wrapping
  if (checked_total_i32(dpos, n, "spadd", &nnzC) != TACO_B200_OK) { device_result_free(dpos); return TACO_B200_ERR_ARG; }
  TB_TRY(exclusive_scan_i32(dpos, dpos, (long long)n + 1));
  int* dcrd = nullptr;
  TB_TRY(device_result_alloc((void**)&dcrd, sizeof(int) * (size_t)(nnzC > 0 ? nnzC : 1)));
  void* dvals = nullptr;
  if (with_vals) TB_TRY(device_result_alloc(&dvals, sizeof(T) * (size_t)(nnzC > 0 ? nnzC : 1)));
  if (n > 0 && nnzC > 0) {
    if (with_vals) {
      ProfScope ps("spadd_numeric");
      TB_TRY((spadd_block_launch<T, 3>(n, s.nnzA, s.nnzB, s.apos.as<int>(), s.acrd.as<int>(), av, s.bpos.as<int>(), s.bcrd.as<int>(),
                                       bv, dpos, dcrd, (T*)dvals, nullptr)));
    } else {
      TB_TRY((spadd_block_launch<T, 1>(n, s.nnzA, s.nnzB, s.apos.as<int>(), s.acrd.as<int>(), nullptr, s.bpos.as<int>(),
                                       s.bcrd.as<int>(), nullptr, dpos, dcrd, nullptr, nullptr)));
    }
  }
  return publish_structure(C, n, dpos, dcrd, nnzC, sizeof(T), dvals);
}

template <typename T>
static int spadd_numeric(Csr3& s, const In& av, const In& bv, const In& cpos, Out& cv) {
  const int n = s.A.rows;
  if (n > 0) {
    ProfScope ps("spadd_numeric");
    TB_TRY((spadd_block_launch<T, 2>(n, s.nnzA, s.nnzB, s.apos.as<int>(), s.acrd.as<int>(), av.as<T>(), s.bpos.as<int>(),
                                     s.bcrd.as<int>(), bv.as<T>(), cpos.as<int>(), nullptr, cv.as<T>(), nullptr)));
  }
  return TACO_B200_OK;
}

// ---- SpGEMM host side ---------------------------------------------------------------------------------------------
struct SgBins {
  void* bin_count = nullptr;
  void* bin_rows = nullptr;
  void* bitmaps = nullptr;
  int h[SG_BINS] = {0, 0, 0, 0, 0};
  int big_grid = 0;
  const int* rows(int b, int n) const { return (const int*)bin_rows + (size_t)b * n; }
  ~SgBins() { scratch_free(bin_count); scratch_free(bin_rows); scratch_free(bitmaps); }
};

// bin the rows by their number of products (and store 0 into counts[] for rows without any, if counts != nullptr)
static int spgemm_make_bins(Csr3& s, int* counts, SgBins* b) {
  const int n = s.A.rows;
  TB_TRY(scratch_alloc(&b->bin_count, sizeof(int) * 8));
  TB_TRY(scratch_alloc(&b->bin_rows, sizeof(int) * SG_BINS * (size_t)(n > 0 ? n : 1)));
  TB_CUDA(cudaMemsetAsync(b->bin_count, 0, sizeof(int) * 8, stream()));
  if (n > 0) {
    spgemm_bound_kernel<<<(n + 255) / 256, 256, 0, stream()>>>(n, s.apos.as<int>(), s.acrd.as<int>(), s.bpos.as<int>(),
                                                              (int*)b->bin_count, (int*)b->bin_rows, counts);
    count_launch(1);
    TB_TRY(read_back(b->h, b->bin_count, sizeof(int) * SG_BINS));
  }
  if (b->h[4] > 0) {
    b->big_grid = b->h[4] < 2 * num_sms() ? b->h[4] : 2 * num_sms();
    TB_TRY(scratch_alloc(&b->bitmaps, sizeof(unsigned) * (size_t)b->big_grid * ((s.B.cols + 31) / 32)));
  }
  return TACO_B200_OK;
}

static int spgemm_group_lanes(const Csr3& s) {   // lanes that share one A entry in the warp kernels
  const double avg = s.B.rows > 0 ? (double)s.nnzB / s.B.rows : 0.0;
  return avg <= 8.0 ? 8 : (avg <= 16.0 ? 16 : 32);
}

template <typename T, typename KEY, bool CRD, bool VALS>
static void spgemm_fill_warp_bins(const SgBins& b, Csr3& s, const T* av, const T* bv, const int* dpos, int* dcrd, T* dvals) {
  const int n = s.A.rows, G = spgemm_group_lanes(s);
#define TB_SG_FILL(BIN, EPL)                                                                                            \
  if (b.h[BIN] > 0) {                                                                                                   \
    spgemm_fill_warp_kernel<T, KEY, EPL, CRD, VALS><<<(b.h[BIN] + 7) / 8, 256, 0, stream()>>>(                          \
        b.rows(BIN, n), b.h[BIN], G, s.apos.as<int>(), s.acrd.as<int>(), av, s.bpos.as<int>(), s.bcrd.as<int>(), bv, dpos, dcrd, \
        dvals);                                                                                                         \
    count_launch(1);                                                                                                    \
  }
  TB_SG_FILL(0, 2)
  TB_SG_FILL(1, 4)
  TB_SG_FILL(2, 8)
#undef TB_SG_FILL
}

template <typename T>
static void spgemm_numeric_rows(const int* rows_list, int nrows, Csr3& s, const T* av, const T* bv, const int* cpos,
                                const int* ccrd, T* cv) {
  if (nrows <= 0) return;
  const double avg = s.B.rows > 0 ? (double)s.nnzB / s.B.rows : 0.0;
#define TB_SGN(TEAM)                                                                                                   \
  spgemm_numeric_kernel<T, TEAM><<<(unsigned)(((long long)nrows * TEAM + 255) / 256), 256, 0, stream()>>>(              \
      rows_list, nrows, s.apos.as<int>(), s.acrd.as<int>(), av, s.bpos.as<int>(), s.bcrd.as<int>(), bv, cpos, ccrd, cv)
  if (avg <= 12.0) TB_SGN(8);
  else if (avg <= 24.0) TB_SGN(16);
  else TB_SGN(32);
#undef TB_SGN
  count_launch(1);
}

// symbolic count -> scan -> (host learns nnz) -> fill.  av != nullptr: the fill also computes the values (evaluate).
template <typename T>
static int spgemm_assemble_impl(taco_tensor_t* C, Csr3& s, const T* av, const T* bv) {
  const int n = s.A.rows, ncols = s.B.cols;
  const bool with_vals = av != nullptr;
  const bool key32 = ncols <= (1 << 24);
  int* dpos = nullptr;
  TB_TRY(device_result_alloc((void**)&dpos, sizeof(int) * ((size_t)n + 1)));
  TB_CUDA(cudaMemsetAsync(dpos + n, 0, sizeof(int), stream()));
  SgBins b;
  TB_TRY(spgemm_make_bins(s, dpos, &b));
  const int G = spgemm_group_lanes(s);
  {
    ProfScope ps("spgemm_symbolic");
#define TB_SG_COUNT(BIN, EPL)                                                                                            \
    if (b.h[BIN] > 0) {                                                                                                  \
      spgemm_count_warp_kernel<EPL><<<(b.h[BIN] + 7) / 8, 256, 0, stream()>>>(b.rows(BIN, n), b.h[BIN], G, s.apos.as<int>(), \
          s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), dpos);                                                   \
      count_launch(1);                                                                                                   \
    }
    TB_SG_COUNT(0, 2)
    TB_SG_COUNT(1, 4)
    TB_SG_COUNT(2, 8)
#undef TB_SG_COUNT
    if (b.h[3] > 0) {
      spgemm_symbolic_kernel<256, SG_CTA_CAP, false><<<b.h[3], 256, 0, stream()>>>(b.rows(3, n), b.h[3], s.apos.as<int>(),
          s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), dpos, dpos, nullptr);
      count_launch(1);
    }
    if (b.h[4] > 0) {
      spgemm_symbolic_bitmap_kernel<false><<<b.big_grid, 256, 0, stream()>>>(b.rows(4, n), b.h[4], ncols, s.apos.as<int>(),
          s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), (unsigned*)b.bitmaps, dpos, dpos, nullptr);
      count_launch(1);
    }
  }
  TB_CUDA(cudaGetLastError());
  int32_t nnzC = 0;             // 64-bit total of the row counts first: refuses results beyond int32 instead of wrapping
  if (checked_total_i32(dpos, n, "spgemm", &nnzC) != TACO_B200_OK) { device_result_free(dpos); return TACO_B200_ERR_ARG; }
  TB_TRY(exclusive_scan_i32(dpos, dpos, (long long)n + 1));
  int* dcrd = nullptr;
  TB_TRY(device_result_alloc((void**)&dcrd, sizeof(int) * (size_t)(nnzC > 0 ? nnzC : 1)));
  void* dvals = nullptr;
  if (with_vals) TB_TRY(device_result_alloc(&dvals, sizeof(T) * (size_t)(nnzC > 0 ? nnzC : 1)));
  {
    ProfScope ps("spgemm_numeric");
    if (with_vals) {
      if (key32) spgemm_fill_warp_bins<T, uint32_t, true, true>(b, s, av, bv, dpos, dcrd, (T*)dvals);
      else spgemm_fill_warp_bins<T, uint64_t, true, true>(b, s, av, bv, dpos, dcrd, (T*)dvals);
    } else {
      if (key32) spgemm_fill_warp_bins<T, uint32_t, true, false>(b, s, av, bv, dpos, dcrd, (T*)dvals);
      else spgemm_fill_warp_bins<T, uint64_t, true, false>(b, s, av, bv, dpos, dcrd, (T*)dvals);
    }
    if (b.h[3] > 0) {
      spgemm_symbolic_kernel<256, SG_CTA_CAP, true><<<b.h[3], 256, 0, stream()>>>(b.rows(3, n), b.h[3], s.apos.as<int>(),
          s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), dpos, dpos, dcrd);
      count_launch(1);
      if (with_vals) spgemm_numeric_rows<T>(b.rows(3, n), b.h[3], s, av, bv, dpos, dcrd, (T*)dvals);
    }
    if (b.h[4] > 0) {
      spgemm_symbolic_bitmap_kernel<true><<<b.big_grid, 256, 0, stream()>>>(b.rows(4, n), b.h[4], ncols, s.apos.as<int>(),
          s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), (unsigned*)b.bitmaps, dpos, dpos, dcrd);
      count_launch(1);
      if (with_vals) spgemm_numeric_rows<T>(b.rows(4, n), b.h[4], s, av, bv, dpos, dcrd, (T*)dvals);
    }
  }
  TB_CUDA(cudaGetLastError());
  return publish_structure(C, n, dpos, dcrd, nnzC, sizeof(T), dvals);
}

// compute() on an assembled result: rows with few products re-derive their sorted order in registers and write only
// the values; long rows locate each product in the row's crd by binary search.
template <typename T>
static int spgemm_numeric(Csr3& s, const In& av, const In& bv, const In& cpos, const In& ccrd, Out& cv) {
  const int n = s.A.rows;
  if (n == 0) return TACO_B200_OK;
  SgBins b;
  TB_TRY(spgemm_make_bins(s, nullptr, &b));
  ProfScope ps("spgemm_numeric");
  if (s.B.cols <= (1 << 24))
    spgemm_fill_warp_bins<T, uint32_t, false, true>(b, s, av.as<T>(), bv.as<T>(), cpos.as<int>(), nullptr, cv.as<T>());
  else
    spgemm_fill_warp_bins<T, uint64_t, false, true>(b, s, av.as<T>(), bv.as<T>(), cpos.as<int>(), nullptr, cv.as<T>());
  spgemm_numeric_rows<T>(b.rows(3, n), b.h[3], s, av.as<T>(), bv.as<T>(), cpos.as<int>(), ccrd.as<int>(), cv.as<T>());
  spgemm_numeric_rows<T>(b.rows(4, n), b.h[4], s, av.as<T>(), bv.as<T>(), cpos.as<int>(), ccrd.as<int>(), cv.as<T>());
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

static int sparse_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, bool product) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, product, &s));
  if (!s.C.pos) return fail(TACO_B200_ERR_ARG, "result has no structure: call assemble first");
  int32_t nnzC = 0;
  TB_TRY(csr_nnz(s.C, C->vals_size, &nnzC));
  size_t es = dsize(s.A.dt);
  In av, bv, cpos, ccrd; Out cv;
  TB_TRY(av.acquire(s.A.vals ? s.A.vals : (void*)s.A.pos, es * (size_t)s.nnzA));
  TB_TRY(bv.acquire(s.B.vals ? s.B.vals : (void*)s.B.pos, es * (size_t)s.nnzB));
  TB_TRY(cpos.acquire(s.C.pos, sizeof(int32_t) * ((size_t)s.C.rows + 1)));
  if (nnzC > 0) {
    TB_TRY(cv.acquire(s.C.vals, es * (size_t)nnzC));
    if (product) {
      TB_TRY(ccrd.acquire(s.C.crd, sizeof(int32_t) * (size_t)nnzC));
      if (s.A.dt == DType::F64) TB_TRY(spgemm_numeric<double>(s, av, bv, cpos, ccrd, cv));
      else TB_TRY(spgemm_numeric<float>(s, av, bv, cpos, ccrd, cv));
    } else {
      if (s.A.dt == DType::F64) TB_TRY(spadd_numeric<double>(s, av, bv, cpos, cv));
      else TB_TRY(spadd_numeric<float>(s, av, bv, cpos, cv));
    }
    TB_TRY(cv.commit());
  }
  return finish_call();
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_spadd_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, false, &s));
  if (s.A.dt == DType::F64) return spadd_assemble_impl<double>(C, s, nullptr, nullptr);
  return spadd_assemble_impl<float>(C, s, nullptr, nullptr);
}
int taco_b200_spadd_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) { return sparse_compute(C, A, B, false); }
// evaluate = assemble + compute with the value fill fused into the structure fill (one merge pass less)
int taco_b200_spadd_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, false, &s));
  const size_t es = dsize(s.A.dt);
  In av, bv;
  TB_TRY(av.acquire(s.A.vals ? s.A.vals : (void*)s.A.pos, es * (size_t)s.nnzA));
  TB_TRY(bv.acquire(s.B.vals ? s.B.vals : (void*)s.B.pos, es * (size_t)s.nnzB));
  if (s.A.dt == DType::F64) TB_TRY(spadd_assemble_impl<double>(C, s, av.as<double>(), bv.as<double>()));
  else TB_TRY(spadd_assemble_impl<float>(C, s, av.as<float>(), bv.as<float>()));
  return finish_call();
}

int taco_b200_spgemm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, true, &s));
  if (s.A.dt == DType::F64) return spgemm_assemble_impl<double>(C, s, nullptr, nullptr);
  return spgemm_assemble_impl<float>(C, s, nullptr, nullptr);
}
int taco_b200_spgemm_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) { return sparse_compute(C, A, B, true); }
// evaluate = assemble + compute with the values produced by the same sort that orders the columns
int taco_b200_spgemm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, true, &s));
  const size_t es = dsize(s.A.dt);
  In av, bv;
  TB_TRY(av.acquire(s.A.vals ? s.A.vals : (void*)s.A.pos, es * (size_t)s.nnzA));
  TB_TRY(bv.acquire(s.B.vals ? s.B.vals : (void*)s.B.pos, es * (size_t)s.nnzB));
  if (s.A.dt == DType::F64) TB_TRY(spgemm_assemble_impl<double>(C, s, av.as<double>(), bv.as<double>()));
  else TB_TRY(spgemm_assemble_impl<float>(C, s, av.as<float>(), bv.as<float>()));
  return finish_call();
}

}  // extern "C"
