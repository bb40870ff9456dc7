// spmm.cu -- C(i,k) = A(i,j) * B(j,k), A CSR, B dense row-major, C dense (row-major or {1,0}), fp32 / fp64.
//
// Replaces the CUDA the reference emits for scheduleSpMMGPU (/root/reference/test/tests-scheduling-eval.cpp:249-268,
// SURVEY.md Appendix A.2):
//   reference: grid ceil(nnz/64), warp gets 8 nnz, lane <-> k%32, the binary search is redone for each of 4
//              `dense_val` passes, ONE GLOBAL atomicAdd PER (nnz, k) (8.2 G atomics at config C2), host-serial
//              zeroing of the 2 GB result in managed memory, K <= 128 only.
//   here     : nnz-balanced ROW-ALIGNED slots: warp w owns the rows whose first nonzero lies in [w*W,(w+1)*W)
//              (one binary search per slot in a pre-pass, the search of taco_binarySearchBeforeBlock,
//              /root/reference/src/codegen/codegen_cuda.cpp:110-125).  A lane owns 16 bytes of the dense row
//              (4 fp32 / 2 fp64 columns), so every gathered row of B is one fully coalesced 512-byte warp load
//              (ld.global.nc, L1-allocating: hot columns of a power-law matrix stay in L1/L2), 8 independent row
//              gathers are in flight per warp, accumulators live in registers and each C row is written exactly
//              once with a streaming 128-bit store -- no zero-fill pass, no atomics.  Empty rows are zeroed by
//              their owner.  Only "hub" rows (longer than LONG nonzeros) are split across the slots they span;
//              their partial sums are combined with vector red.global.add into a pre-zeroed row.
//              Inside a row the products are accumulated in ascending position order with separate multiply and
//              add -- the reference C kernel's order (Appendix A.1) -- so non-hub rows are bit-identical to it.
// Algorithmic bytes per launch: nnz*(4+sizeof T) + 4(n+1) + sizeof T*K*(cols + rows).
#include <cstdlib>

#include "common.cuh"

namespace tb {

constexpr int SPMM_W = 64;          // nonzeros per slot (one warp)
constexpr int SPMM_LONG = 512;      // rows longer than this are split across slots

template <typename T, int VEC> struct Frag { T v[VEC]; };

// A lane's 16-byte piece of a gathered B row.  L1-allocating (hot columns of a power-law matrix are re-used inside an
// SM) and tagged evict_last in L2: the dense operand is the only array with re-use, the CSR arrays and C stream by.
template <typename T, int VEC>
__device__ __forceinline__ Frag<T, VEC> load_row(const T* __restrict__ p, uint64_t keep) {
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

template <typename T, int VEC, bool COLMAJOR>
__device__ __forceinline__ void store_row(T* __restrict__ C, size_t row, int col, int rows, int K, const Frag<T, VEC>& f,
                                          uint64_t strm) {
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

template <typename T, int VEC, bool COLMAJOR>
__device__ __forceinline__ void red_row(T* __restrict__ C, size_t row, int col, int rows, int K, const Frag<T, VEC>& f) {
  if constexpr (COLMAJOR) {
#pragma unroll
    for (int e = 0; e < VEC; e++) atomicAdd(C + (size_t)(col + e) * rows + row, f.v[e]);
  } else {
    T* p = C + row * K + col;
    if constexpr (VEC == 4 && sizeof(T) == 4) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(f.v[0]), "f"(f.v[1]), "f"(f.v[2]),
                   "f"(f.v[3]) : "memory");
    } else {
#pragma unroll
      for (int e = 0; e < VEC; e++) atomicAdd(p + e, f.v[e]);
    }
  }
}

// A launch covers the row range [r0, r1) = nonzeros [p0, p1) (the whole matrix, or one row chunk of the host-operand
// pipeline below).
struct SpmmRange { int r0, r1, p0, p1; };

// Pre-pass: slot_rows[w] = first row whose first nonzero is at or after p0 + w*W; slot_rows[nslots] = r1.
// Also zeroes the C row of every hub row (done by the slot in which the hub row's first slot boundary falls).
// RMAP (here and in spmm_csr_kernel): result row r is stored at row rowmap[r] of C (TTM: the rows are the fibers of a CSF
// tensor, rowmap their cells in the dense (i,j) plane, csf.cu); the unmapped instantiations are unchanged by the flag.
template <typename T, bool COLMAJOR, bool RMAP = false>
__global__ void spmm_slot_rows_kernel(const int* __restrict__ pos, SpmmRange rg, int rows, int nslots, int K,
                                      int* __restrict__ slot_rows, T* __restrict__ C, const unsigned* __restrict__ rowmap = nullptr) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nslots) return;
  if (w == nslots) { slot_rows[w] = rg.r1; return; }
  int lo = rg.p0 + w * SPMM_W;
  slot_rows[w] = tbd::search_first_ge(pos, rg.r0, rg.r1, lo);
  if (lo < rg.p1) {
    int i = tbd::search_last_le(pos, rg.r0, rg.r1, lo);      // the row that contains nonzero `lo`
    int s = __ldg(pos + i), e = __ldg(pos + i + 1);
    if (e - s > SPMM_LONG && lo - s < SPMM_W) {
      const size_t orow = RMAP ? (size_t)__ldg(rowmap + i) : (size_t)i;
      if constexpr (COLMAJOR) { for (int k = 0; k < K; k++) C[(size_t)k * rows + orow] = T(0); }
      else { for (int k = 0; k < K; k++) C[orow * K + k] = T(0); }
    }
  }
}

// acc += sum over nonzeros p in [a,b) of vals[p] * B[crd[p], col..col+VEC), in ascending p with separate multiply and
// add (the reference C kernel's order, Appendix A.1).  32 nonzeros are fetched with one coalesced load per array and
// broadcast by shuffle; U independent B-row gathers are in flight per warp.
template <typename T, int VEC, int U>
__device__ __forceinline__ void spmm_accumulate(Frag<T, VEC>& acc, const int* __restrict__ crd, const T* __restrict__ vals,
                                                const T* __restrict__ Bcol, int K, int a, int b, int lane, uint64_t keep,
                                                uint64_t strm) {
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
        bv[u] = load_row<T, VEC>(Bcol + (size_t)c * K, keep);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const T v = __shfl_sync(0xffffffffu, my_v, j + u);
#pragma unroll
        for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + v * bv[u].v[x];   // mul then add: never fused
      }
    }
    if (j < cnt) {                      // 1 .. U-1 left: same two phases, warp-uniform predicates
      const int rem = cnt - j;
      Frag<T, VEC> bv[U > 1 ? U - 1 : 1];
#pragma unroll
      for (int u = 0; u < U - 1; u++) {
        if (u < rem) {
          const int c = __shfl_sync(0xffffffffu, my_c, j + u);
          bv[u] = load_row<T, VEC>(Bcol + (size_t)c * K, keep);
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

template <typename T, int VEC, bool COLMAJOR, int U, int WARPS, int MINB, bool RMAP = false>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_csr_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                const T* __restrict__ B, T* __restrict__ C, int rows, int K, SpmmRange rg, int nslots,
                const int* __restrict__ slot_rows, const unsigned* __restrict__ rowmap = nullptr) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (w >= nslots) return;
  const int nnz = rg.p1;
  const int lo = rg.p0 + w * SPMM_W, hi = min(lo + SPMM_W, nnz);
  // the slot's own window of crd / vals is needed a few dependent loads from now: pull it into L2 meanwhile
  if (lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(crd + lo + lane * 32));
  else if (lane < 2 + (int)(2 * sizeof(T) / 4)) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + lo + (lane - 2) * (128 / (int)sizeof(T))));
  const uint64_t keep = 0, strm = 0;     // (L2 eviction-policy hints measured no change in hit rate at C2: not used)
  const int col = (blockIdx.y * 32 + lane) * VEC;
  const bool active = col < K;
  const T* Bcol = B + (active ? col : 0);        // inactive lanes (ragged K) gather column 0 and never store
  const int R0 = __ldg(slot_rows + w), R1 = __ldg(slot_rows + w + 1);

  // Work items of the slot, all run through ONE copy of the accumulate loop:
  //   group -1 (optional): the piece [lo, min(hi,e)) of a hub row that started in an earlier slot (added atomically),
  //   groups 0..: the rows this slot owns, 32 at a time -- empty rows are zeroed, every other row is accumulated in
  //   registers and written exactly once; a hub row (always the last row a slot owns) contributes its first piece
  //   atomically into the row the pre-pass zeroed.
  int tail_e = 0;
  bool tail = false;
  if (R0 > rg.r0 && lo < nnz) {
    const int s = __ldg(pos + R0 - 1);
    tail_e = __ldg(pos + R0);
    tail = tail_e > lo && tail_e - s > SPMM_LONG;
  }
  for (int rb = tail ? R0 - 32 : R0; rb < R1; rb += 32) {
    const bool is_tail = rb < R0;
    const int r = rb + lane;
    const bool valid = !is_tail && r < R1;
    int s = valid ? __ldg(pos + r) : 0;
    int e = valid ? __ldg(pos + r + 1) : 0;
    unsigned empty = __ballot_sync(0xffffffffu, valid && e == s);
    unsigned full = __ballot_sync(0xffffffffu, valid && e > s);
    int rowbase = rb;
    if (is_tail) { s = lo; e = min(hi, tail_e); full = 1u; rowbase = R0 - 1; }
    Frag<T, VEC> acc;
#pragma unroll
    for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
    while (empty) {
      const int h = __ffs(empty) - 1;
      empty &= empty - 1;
      if (active) store_row<T, VEC, COLMAJOR>(C, RMAP ? (size_t)__ldg(rowmap + rb + h) : (size_t)(rb + h), col, rows, K, acc, strm);
    }
    while (full) {
      const int h = __ffs(full) - 1;
      full &= full - 1;
      const int hs = __shfl_sync(0xffffffffu, s, h);
      int he = __shfl_sync(0xffffffffu, e, h);
      const bool hub = is_tail || he - hs > SPMM_LONG;
      if (hub) he = min(hi, he);
#pragma unroll
      for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
      spmm_accumulate<T, VEC, U>(acc, crd, vals, Bcol, K, hs, he, lane, keep, strm);
      if (active) {
        const size_t orow = RMAP ? (size_t)__ldg(rowmap + rowbase + h) : (size_t)(rowbase + h);
        if (hub) red_row<T, VEC, COLMAJOR>(C, orow, col, rows, K, acc);
        else store_row<T, VEC, COLMAJOR>(C, orow, col, rows, K, acc, strm);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Ring kernel (row-major C, 16-byte lane fragments): the same slot/ownership scheme as spmm_csr_kernel, but the B-row
// gathers no longer land in registers.  Every lane copies ITS 16 bytes of a gathered row with cp.async (LDGSTS) into a
// per-warp ring of D rows in shared memory and later reads the same 16 bytes back, so no cross-lane synchronisation is
// needed and D gathers stay in flight per warp at no register cost.  The nonzeros of all non-hub rows a slot owns form
// ONE flat stream [pos[R0], pos[R1e)): the producer runs D nonzeros ahead of the consumer ACROSS row boundaries, which
// removes the per-row latency chain (crd -> B row -> store) that bounded the register kernel on short rows (measured:
// halving the gathered bytes per pass only cut its time by 30 %).  Within a row the products are still accumulated in
// ascending position order with separate multiply and add, so non-hub rows stay bit-identical to the reference.
// ---------------------------------------------------------------------------------------------------------
template <bool CA>
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  if constexpr (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int VEC>
__device__ __forceinline__ Frag<T, VEC> lds_frag(uint32_t addr) {
  Frag<T, VEC> f;
  if constexpr (sizeof(T) == 4) {
    static_assert(VEC == 4, "ring kernel: 16-byte fragments");
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(f.v[0]), "=f"(f.v[1]), "=f"(f.v[2]), "=f"(f.v[3]) : "r"(addr) : "memory");
  } else {
    static_assert(VEC == 2, "ring kernel: 16-byte fragments");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(f.v[0]), "=d"(f.v[1]) : "r"(addr) : "memory");
  }
  return f;
}

template <typename T, int VEC, int D, int WARPS, int MINB, bool CA>
__global__ void __launch_bounds__(WARPS * 32, MINB)
spmm_ring_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                 const T* __restrict__ B, T* __restrict__ C, int rows, int K, SpmmRange rg, int nslots,
                 const int* __restrict__ slot_rows) {
  static_assert(D >= 2 && D <= 32 && (D & (D - 1)) == 0, "ring depth: power of two, at most one chunk");
  extern __shared__ __align__(16) unsigned char ring_raw[];
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (w >= nslots) return;
  const int nnz = rg.p1;
  const int lo = rg.p0 + w * SPMM_W, hi = min(lo + SPMM_W, nnz);
  if (lane < 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(crd + lo + lane * 32));
  else if (lane < 2 + (int)(2 * sizeof(T) / 4)) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals + lo + (lane - 2) * (128 / (int)sizeof(T))));
  const int col = (blockIdx.y * 32 + lane) * VEC;
  const bool active = col < K;
  const T* Bcol = B + (active ? col : 0);        // inactive lanes (ragged K) gather column 0 and never store
  const int R0 = __ldg(slot_rows + w), R1 = __ldg(slot_rows + w + 1);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ring_raw) + (uint32_t)(threadIdx.x >> 5) * (D * 512) + lane * 16;
  Frag<T, VEC> acc;

  // (a) the piece [lo, min(hi, e)) of a hub row that started in an earlier slot: added atomically (cold path)
  if (R0 > rg.r0 && lo < nnz) {
    const int s = __ldg(pos + R0 - 1), e = __ldg(pos + R0);
    if (e > lo && e - s > SPMM_LONG) {
#pragma unroll
      for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
      spmm_accumulate<T, VEC, 1>(acc, crd, vals, Bcol, K, lo, min(hi, e), lane, 0, 0);
      if (active) red_row<T, VEC, false>(C, R0 - 1, col, rows, K, acc);
    }
  }
  if (R1 <= R0) return;
  // (b) empty rows are zeroed by their owner; a hub row (always the last row a slot owns) contributes its first piece
#pragma unroll
  for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
  for (int rb = R0; rb < R1; rb += 32) {
    const int r = rb + lane;
    const bool valid = r < R1;
    const int s = valid ? __ldg(pos + r) : 0, e = valid ? __ldg(pos + r + 1) : 1;
    unsigned empty = __ballot_sync(0xffffffffu, valid && e == s);
    while (empty) {
      const int h = __ffs(empty) - 1;
      empty &= empty - 1;
      if (active) store_row<T, VEC, false>(C, rb + h, col, rows, K, acc, 0);
    }
  }
  int R1e = R1;
  {
    const int s = __ldg(pos + R1 - 1), e = __ldg(pos + R1);
    if (e - s > SPMM_LONG) {
      R1e = R1 - 1;
      spmm_accumulate<T, VEC, 1>(acc, crd, vals, Bcol, K, s, min(hi, e), lane, 0, 0);
      if (active) red_row<T, VEC, false>(C, R1e, col, rows, K, acc);
#pragma unroll
      for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
    }
  }
  if (R1e <= R0) return;
  // (c) the flat stream of the complete rows R0 .. R1e-1
  const int S = __ldg(pos + R0), E = __ldg(pos + R1e);
  if (E <= S) return;
  int crd_k = 0, crd_n = 0;                      // column ids of the chunk being consumed / the next chunk (lane = offset)
  T val_k = T(0), val_n = T(0);
  if (S + lane < E) { crd_k = tbd::ldg_stream_i32(crd + S + lane); val_k = __ldg(vals + S + lane); }
  if (S + 32 + lane < E) { crd_n = tbd::ldg_stream_i32(crd + S + 32 + lane); val_n = __ldg(vals + S + 32 + lane); }
  // prologue: the first D gathers
#pragma unroll
  for (int d = 0; d < D; d++) {
    const int c = __shfl_sync(0xffffffffu, crd_k, d);
    if (S + d < E) cp_async16<CA>(ring + d * 512, Bcol + (size_t)c * K);
    cp_async_commit();
  }
  // row cursor: the ends of 32 rows at a time live in `my_e` (lane = row - rb); rows past R1e read as E
  int rb = R0, row = R0;
  int my_e = (rb + lane < R1e) ? __ldg(pos + rb + lane + 1) : E;
  int row_end = __shfl_sync(0xffffffffu, my_e, 0);
  while (row_end == S) {                          // leading empty rows (already zeroed)
    row++;
    if (row - rb == 32) { rb += 32; my_e = (rb + lane < R1e) ? __ldg(pos + rb + lane + 1) : E; }
    row_end = __shfl_sync(0xffffffffu, my_e, row - rb);
  }
  const char* Bbytes = (const char*)Bcol;
  const unsigned stride = (unsigned)K * (unsigned)sizeof(T);          // bytes between rows of B
  constexpr uint32_t RMASK = D * 512 - 1;
  constexpr int G = D >= 4 ? 4 : 2;                                   // nonzeros per fast-path step
  for (int base = S; base < E; base += 32) {
    const int cnt = min(32, E - base);
    int j = 0;
    while (j < cnt) {
      const int pc = base + j;
      const uint32_t off = ((uint32_t)(pc - S) * 512u) & RMASK;
      const int run = min(row_end - pc, cnt - j);                     // nonzeros left in this row and this chunk
      if (run >= G && pc + D + G <= E && (((j + D) ^ (j + D + G - 1)) & 32) == 0) {
        // fast path: G nonzeros of one row; their G refills are all valid and come from one chunk register
        const int src = (j + D) < 32 ? crd_k : crd_n;
        cp_async_wait<D - G>();
        Frag<T, VEC> bv[G];
        T v[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          bv[g] = lds_frag<T, VEC>(ring + ((off + g * 512u) & RMASK));
          v[g] = __shfl_sync(0xffffffffu, val_k, j + g);
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
#pragma unroll
          for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + v[g] * bv[g].v[x];      // mul then add: never fused
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
          const unsigned c = (unsigned)__shfl_sync(0xffffffffu, src, j + D + g);   // source lane is taken modulo 32
          cp_async16<CA>(ring + ((off + g * 512u) & RMASK), Bbytes + (size_t)c * stride);
          cp_async_commit();
        }
        j += G;
      } else {
        const uint32_t slot = ring + off;
        cp_async_wait<D - 1>();
        const Frag<T, VEC> b1 = lds_frag<T, VEC>(slot);
        const T v1 = __shfl_sync(0xffffffffu, val_k, j);
#pragma unroll
        for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + v1 * b1.v[x];
        const int jn = j + D;
        const unsigned c = (unsigned)__shfl_sync(0xffffffffu, jn < 32 ? crd_k : crd_n, jn);
        if (pc + D < E) cp_async16<CA>(slot, Bbytes + (size_t)c * stride);
        cp_async_commit();
        j += 1;
      }
      if (base + j == row_end) {
        if (active) store_row<T, VEC, false>(C, row, col, rows, K, acc, 0);
#pragma unroll
        for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
        do {                                      // next non-empty row (rows past R1e read as E: ends with base + j == E)
          row++;
          if (row - rb == 32) { rb += 32; my_e = (rb + lane < R1e) ? __ldg(pos + rb + lane + 1) : E; }
          row_end = __shfl_sync(0xffffffffu, my_e, row - rb);
        } while (row_end == base + j && row < R1e);
      }
    }
    crd_k = crd_n; val_k = val_n;
    const int nb = base + 64 + lane;
    if (nb < E) { crd_n = tbd::ldg_stream_i32(crd + nb); val_n = __ldg(vals + nb); }
  }
  cp_async_wait<0>();
}

// Launch variants: (gathers in flight per warp, warps per CTA, min CTAs per SM).  TACO_B200_SPMM_VARIANT selects one
// for tuning runs; the default is the measured best at config C2 (profiles/).
template <typename T, int VEC, bool COLMAJOR, int U, int WARPS, int MINB>
static void spmm_go(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, SpmmRange rg, int nslots,
                    const int* slot_rows, cudaStream_t st) {
  dim3 grid((nslots + WARPS - 1) / WARPS, (K + 32 * VEC - 1) / (32 * VEC));
  spmm_csr_kernel<T, VEC, COLMAJOR, U, WARPS, MINB><<<grid, WARPS * 32, 0, st>>>(pos, crd, vals, B, C, rows, K, rg, nslots,
                                                                                 slot_rows);
}

template <typename T, int VEC, int D, int WARPS, int MINB, bool CA>
static int spmm_ring_go(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, SpmmRange rg,
                        int nslots, const int* slot_rows, cudaStream_t st) {
  constexpr int smem = WARPS * D * 512;
  static bool configured = false;
  if (!configured) {
    TB_CUDA(cudaFuncSetAttribute(spmm_ring_kernel<T, VEC, D, WARPS, MINB, CA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TB_CUDA(cudaFuncSetAttribute(spmm_ring_kernel<T, VEC, D, WARPS, MINB, CA>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
    configured = true;
  }
  dim3 grid((nslots + WARPS - 1) / WARPS, (K + 32 * VEC - 1) / (32 * VEC));
  spmm_ring_kernel<T, VEC, D, WARPS, MINB, CA><<<grid, WARPS * 32, smem, st>>>(pos, crd, vals, B, C, rows, K, rg, nslots, slot_rows);
  return TACO_B200_OK;
}

template <typename T, int VEC, bool COLMAJOR>
static int spmm_launch_impl(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, SpmmRange rg) {
  const int nnz = rg.p1 - rg.p0;
  int nslots = nnz > 0 ? (nnz + SPMM_W - 1) / SPMM_W : 1;
  void* slot_rows = nullptr;
  TB_TRY(scratch_alloc(&slot_rows, sizeof(int) * (size_t)(nslots + 1)));
  spmm_slot_rows_kernel<T, COLMAJOR><<<(nslots + 1 + 255) / 256, 256, 0, stream()>>>(pos, rg, rows, nslots, K,
                                                                                     (int*)slot_rows, C);
  static const int variant = getenv("TACO_B200_SPMM_VARIANT") ? atoi(getenv("TACO_B200_SPMM_VARIANT")) : 0;
  {
    ProfScope ps("spmm_csr");
    const int* sr = (const int*)slot_rows;
    cudaStream_t st = stream();
    constexpr bool RING_OK = !COLMAJOR && VEC * sizeof(T) == 16;
    if constexpr (RING_OK) {
      switch (variant) {
        case 10: TB_TRY((spmm_ring_go<T, VEC, 8, 8, 6, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 11: TB_TRY((spmm_ring_go<T, VEC, 16, 8, 3, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 12: TB_TRY((spmm_ring_go<T, VEC, 4, 8, 8, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 13: TB_TRY((spmm_ring_go<T, VEC, 8, 8, 4, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 14: TB_TRY((spmm_ring_go<T, VEC, 8, 4, 12, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 15: TB_TRY((spmm_ring_go<T, VEC, 16, 4, 6, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 16: TB_TRY((spmm_ring_go<T, VEC, 32, 4, 3, false>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 17: TB_TRY((spmm_ring_go<T, VEC, 8, 8, 6, true>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        case 18: TB_TRY((spmm_ring_go<T, VEC, 16, 8, 3, true>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st))); goto launched;
        default: break;
      }
    }
    switch (variant) {
      case 1: spmm_go<T, VEC, COLMAJOR, 4, 8, 4>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st); break;
      case 2: spmm_go<T, VEC, COLMAJOR, 1, 8, 8>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st); break;
      case 3: spmm_go<T, VEC, COLMAJOR, 2, 8, 6>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st); break;
      default: spmm_go<T, VEC, COLMAJOR, 2, 8, 8>(pos, crd, vals, B, C, rows, K, rg, nslots, sr, st); break;
    }
  launched:;
  }
  count_launch(2);
  scratch_free(slot_rows);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

// Row-mapped launch (csf.cu, TTM): C[rowmap[r], :] = sum_p vals[p] * B[crd[p], :] over the rows r of any (pos, crd, vals) level.
template <typename T, int VEC>
static int spmm_mapped_impl(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, int nnz,
                            const unsigned* rowmap, const char* prof_name) {
  const SpmmRange rg{0, rows, 0, nnz};
  const int nslots = nnz > 0 ? (nnz + SPMM_W - 1) / SPMM_W : 1;
  void* slot_rows = nullptr;
  TB_TRY(scratch_alloc(&slot_rows, sizeof(int) * (size_t)(nslots + 1)));
  spmm_slot_rows_kernel<T, false, true><<<(nslots + 1 + 255) / 256, 256, 0, stream()>>>(pos, rg, rows, nslots, K, (int*)slot_rows, C, rowmap);
  {
    ProfScope ps(prof_name);
    constexpr int WARPS = 8;
    dim3 grid((nslots + WARPS - 1) / WARPS, (K + 32 * VEC - 1) / (32 * VEC));
    // (gathers in flight per warp, min CTAs per SM): rows narrower than a full warp fragment (K * sizeof T < 512 bytes) keep
    // fewer bytes in flight per gather, so deeper unrolling is worth its registers there (TACO_B200_TTM_UNROLL=2..5 to compare)
    static const int variant = getenv("TACO_B200_TTM_UNROLL") ? atoi(getenv("TACO_B200_TTM_UNROLL")) : 0;
#define TB_MAPPED_GO(U, MINB)                                                                                              \
  spmm_csr_kernel<T, VEC, false, U, WARPS, MINB, true><<<grid, WARPS * 32, 0, stream()>>>(pos, crd, vals, B, C, rows, K, rg, nslots, \
                                                                                          (const int*)slot_rows, rowmap)
    switch (variant) {
      case 2: TB_MAPPED_GO(4, 6); break;
      case 3: TB_MAPPED_GO(4, 8); break;
      case 4: TB_MAPPED_GO(4, 4); break;
      case 5: TB_MAPPED_GO(8, 4); break;
      default: TB_MAPPED_GO(2, 8); break;
    }
#undef TB_MAPPED_GO
  }
  count_launch(2);
  scratch_free(slot_rows);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

int spmm_mapped(DType dt, const int* pos, const int* crd, const void* vals, const void* B, void* C, int rows, int K, int nnz,
                const unsigned* rowmap, const char* prof_name) {
  const bool a16 = (((uintptr_t)B | (uintptr_t)C) & 15) == 0;
  if (dt == DType::F64) {
    if (a16 && K % 2 == 0) return spmm_mapped_impl<double, 2>(pos, crd, (const double*)vals, (const double*)B, (double*)C, rows, K, nnz, rowmap, prof_name);
    return spmm_mapped_impl<double, 1>(pos, crd, (const double*)vals, (const double*)B, (double*)C, rows, K, nnz, rowmap, prof_name);
  }
  if (a16 && K % 4 == 0) return spmm_mapped_impl<float, 4>(pos, crd, (const float*)vals, (const float*)B, (float*)C, rows, K, nnz, rowmap, prof_name);
  return spmm_mapped_impl<float, 1>(pos, crd, (const float*)vals, (const float*)B, (float*)C, rows, K, nnz, rowmap, prof_name);
}

template <typename T>
static int spmm_launch(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, SpmmRange rg,
                       bool colmajor) {
  constexpr int V = 16 / sizeof(T);
  bool vec_ok = (K % V == 0) && (((uintptr_t)B & 15) == 0) && (((uintptr_t)C & 15) == 0);
  // TACO_B200_SPMM_SLICE=1: one column per lane, so the launch becomes ceil(K/32) passes over A (grid.y, scheduled one
  // after the other), each gathering 32-column slices of the rows of B -- a 4x smaller working set per pass in L2
  static const int sliced = getenv("TACO_B200_SPMM_SLICE") ? atoi(getenv("TACO_B200_SPMM_SLICE")) : 0;
  if (sliced == 1) vec_ok = false;
  if constexpr (sizeof(T) == 4) {          // =2: two columns per lane (64-column slices, 256-byte gathers)
    if (sliced == 2 && vec_ok && !colmajor) return spmm_launch_impl<T, 2, false>(pos, crd, vals, B, C, rows, K, rg);
  }
  if (colmajor) {
    if (vec_ok) return spmm_launch_impl<T, V, true>(pos, crd, vals, B, C, rows, K, rg);
    return spmm_launch_impl<T, 1, true>(pos, crd, vals, B, C, rows, K, rg);
  }
  if (vec_ok) return spmm_launch_impl<T, V, false>(pos, crd, vals, B, C, rows, K, rg);
  return spmm_launch_impl<T, 1, false>(pos, crd, vals, B, C, rows, K, rg);
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
  TB_TRY(scratch_alloc(&dpos, sizeof(int) * ((size_t)rows + 1)));
  TB_TRY(scratch_alloc(&dcrd, sizeof(int) * (size_t)nnz));
  TB_TRY(scratch_alloc(&dvals, es * (size_t)nnz));
  TB_TRY(scratch_alloc(&dC, es * (size_t)rows * K));
  if (b_host) TB_TRY(scratch_alloc(&dB, es * (size_t)Av.cols * K));
  else { TB_TRY(bin.acquire(Bv.vals, es * (size_t)Av.cols * K)); dB = (void*)bin.dptr; }
  cudaEvent_t ready, done_all;
  TB_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  TB_CUDA(cudaEventCreateWithFlags(&done_all, cudaEventDisableTiming));
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
    cudaEvent_t e_up, e_done;
    TB_CUDA(cudaEventCreateWithFlags(&e_up, cudaEventDisableTiming));
    TB_CUDA(cudaEventCreateWithFlags(&e_done, cudaEventDisableTiming));
    if (p1 > p0) {
      TB_CUDA(cudaMemcpyAsync((int*)dcrd + p0, Av.crd + p0, sizeof(int) * (size_t)(p1 - p0), cudaMemcpyHostToDevice, up));
      TB_CUDA(cudaMemcpyAsync((T*)dvals + p0, (const T*)Av.vals + p0, es * (size_t)(p1 - p0), cudaMemcpyHostToDevice, up));
    }
    TB_CUDA(cudaEventRecord(e_up, up));
    TB_CUDA(cudaStreamWaitEvent(main, e_up, 0));
    rc = spmm_launch<T>((const int*)dpos, (const int*)dcrd, (const T*)dvals, (const T*)dB, (T*)dC, rows, K,
                        SpmmRange{r0, r1, p0, p1}, false);
    TB_CUDA(cudaEventRecord(e_done, main));
    TB_CUDA(cudaStreamWaitEvent(down, e_done, 0));
    TB_CUDA(cudaMemcpyAsync((T*)Cv.vals + (size_t)r0 * K, (const T*)dC + (size_t)r0 * K, es * (size_t)(r1 - r0) * K,
                            cudaMemcpyDeviceToHost, down));
    cudaEventDestroy(e_up);
    cudaEventDestroy(e_done);
    r0 = r1;
  }
  TB_CUDA(cudaEventRecord(done_all, down));
  TB_CUDA(cudaStreamWaitEvent(main, done_all, 0));  // frees below (and the caller's sync) are ordered after the downloads
  cudaEventDestroy(ready);
  cudaEventDestroy(done_all);
  scratch_free(dpos); scratch_free(dcrd); scratch_free(dvals); scratch_free(dC);
  if (b_host) scratch_free(dB);
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
                                Av.rows, K, all, cm));
    else
      TB_TRY(spmm_launch<double>(pos.as<int>(), crd.as<int>(), vals.as<double>(), bin.as<double>(), cout.as<double>(),
                                 Av.rows, K, all, cm));
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
  if (classify(Av.pos1) == Mem::Device && A->vals_size > 0) nnz = A->vals_size;
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
      rc = spmm_launch<float>((const int*)pos_full, crd.as<int>(), vals.as<float>(), bin.as<float>(), cout.as<float>(), Av.rows, K, all, cm);
    else
      rc = spmm_launch<double>((const int*)pos_full, crd.as<int>(), vals.as<double>(), bin.as<double>(), cout.as<double>(), Av.rows, K, all, cm);
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
