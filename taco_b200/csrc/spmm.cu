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
#include "common.cuh"

namespace tb {

constexpr int SPMM_W = 64;          // nonzeros per slot (one warp)
constexpr int SPMM_LONG = 512;      // rows longer than this are split across slots
constexpr int SPMM_UNROLL = 8;      // independent B-row gathers in flight per warp
constexpr int SPMM_WARPS = 8;       // warps per CTA

template <typename T, int VEC> struct Frag { T v[VEC]; };

template <typename T, int VEC>
__device__ __forceinline__ Frag<T, VEC> load_row(const T* __restrict__ p) {
  Frag<T, VEC> f;
  if constexpr (VEC == 4 && sizeof(T) == 4) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    f.v[0] = a.x; f.v[1] = a.y; f.v[2] = a.z; f.v[3] = a.w;
  } else if constexpr (VEC == 2 && sizeof(T) == 8) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p));
    f.v[0] = a.x; f.v[1] = a.y;
  } else {
    f.v[0] = __ldg(p);
  }
  return f;
}

template <typename T, int VEC, bool COLMAJOR>
__device__ __forceinline__ void store_row(T* __restrict__ C, int row, int col, int rows, int K, const Frag<T, VEC>& f) {
  if constexpr (COLMAJOR) {
#pragma unroll
    for (int e = 0; e < VEC; e++) C[(size_t)(col + e) * rows + row] = f.v[e];
  } else {
    T* p = C + (size_t)row * K + col;
    if constexpr (VEC == 4 && sizeof(T) == 4) tbd::stg_stream_f4(p, make_float4(f.v[0], f.v[1], f.v[2], f.v[3]));
    else if constexpr (VEC == 2 && sizeof(T) == 8) tbd::stg_stream_d2(p, make_double2(f.v[0], f.v[1]));
    else p[0] = f.v[0];
  }
}

template <typename T, int VEC, bool COLMAJOR>
__device__ __forceinline__ void red_row(T* __restrict__ C, int row, int col, int rows, int K, const Frag<T, VEC>& f) {
  if constexpr (COLMAJOR) {
#pragma unroll
    for (int e = 0; e < VEC; e++) atomicAdd(C + (size_t)(col + e) * rows + row, f.v[e]);
  } else {
    T* p = C + (size_t)row * K + col;
    if constexpr (VEC == 4 && sizeof(T) == 4) {
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(f.v[0]), "f"(f.v[1]), "f"(f.v[2]),
                   "f"(f.v[3]) : "memory");
    } else {
#pragma unroll
      for (int e = 0; e < VEC; e++) atomicAdd(p + e, f.v[e]);
    }
  }
}

// Pre-pass: slot_rows[w] = first row whose first nonzero is at or after w*W; slot_rows[nslots] = rows.
// Also zeroes the C row of every hub row (done by the slot in which the hub row's first slot boundary falls).
template <typename T, bool COLMAJOR>
__global__ void spmm_slot_rows_kernel(const int* __restrict__ pos, int rows, int nnz, int nslots, int K,
                                      int* __restrict__ slot_rows, T* __restrict__ C) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nslots) return;
  if (w == nslots) { slot_rows[w] = rows; return; }
  int lo = w * SPMM_W;
  slot_rows[w] = tbd::search_first_ge(pos, 0, rows, lo);
  if (lo < nnz) {
    int i = tbd::search_last_le(pos, 0, rows, lo);      // the row that contains nonzero `lo`
    int s = __ldg(pos + i), e = __ldg(pos + i + 1);
    if (e - s > SPMM_LONG && lo - s < SPMM_W) {
      if constexpr (COLMAJOR) { for (int k = 0; k < K; k++) C[(size_t)k * rows + i] = T(0); }
      else { for (int k = 0; k < K; k++) C[(size_t)i * K + k] = T(0); }
    }
  }
}

// Accumulate nonzeros [a,b) (all of one row piece, or a run of whole rows) into acc with row tracking.
// ROWS=true : the range is a run of whole short rows rb .. rb+nrows-1 whose end offsets are in register `e` of lane
//             (row - rb); every finished row (including empty ones) is stored.
// ROWS=false: the range is a piece of hub row `rb`; the partial sum is added atomically.
template <typename T, int VEC, bool COLMAJOR, bool ROWS>
__device__ __forceinline__ void spmm_walk(const int* __restrict__ crd, const T* __restrict__ vals,
                                          const T* __restrict__ B, T* __restrict__ C, int rows, int K, int col,
                                          bool active, int lane, int a, int b, int rb, int nrows, int e) {
  Frag<T, VEC> acc;
#pragma unroll
  for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
  int cur = 0;
  int cur_end = ROWS ? __shfl_sync(0xffffffffu, e, 0) : b;
  for (int pb = a; pb < b; pb += 32) {
    const int cnt = min(32, b - pb);
    int my_c = 0;
    T my_v = T(0);
    if (lane < cnt) {
      my_c = tbd::ldg_stream_i32(crd + pb + lane);
      my_v = __ldg(vals + pb + lane);
    }
    for (int j0 = 0; j0 < cnt; j0 += SPMM_UNROLL) {
      Frag<T, VEC> bv[SPMM_UNROLL];
#pragma unroll
      for (int u = 0; u < SPMM_UNROLL; u++) {
        if (j0 + u < cnt) {
          int c = __shfl_sync(0xffffffffu, my_c, j0 + u);
          if (active) bv[u] = load_row<T, VEC>(B + (size_t)c * K + col);
        }
      }
#pragma unroll
      for (int u = 0; u < SPMM_UNROLL; u++) {
        if (j0 + u < cnt) {
          if constexpr (ROWS) {
            const int p = pb + j0 + u;
            while (p == cur_end) {       // warp-uniform: row rb+cur is complete (possibly empty)
              if (active) store_row<T, VEC, COLMAJOR>(C, rb + cur, col, rows, K, acc);
#pragma unroll
              for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
              cur++;
              cur_end = __shfl_sync(0xffffffffu, e, cur);
            }
          }
          T v = __shfl_sync(0xffffffffu, my_v, j0 + u);
          if (active) {
#pragma unroll
            for (int x = 0; x < VEC; x++) acc.v[x] = acc.v[x] + v * bv[u].v[x];   // mul then add: never fused
          }
        }
      }
    }
  }
  if constexpr (ROWS) {
    while (cur < nrows) {                // last row of the run, then trailing empty rows
      if (active) store_row<T, VEC, COLMAJOR>(C, rb + cur, col, rows, K, acc);
#pragma unroll
      for (int x = 0; x < VEC; x++) acc.v[x] = T(0);
      cur++;
    }
  } else {
    if (active) red_row<T, VEC, COLMAJOR>(C, rb, col, rows, K, acc);
  }
}

template <typename T, int VEC, bool COLMAJOR>
__global__ void __launch_bounds__(SPMM_WARPS * 32)
spmm_csr_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                const T* __restrict__ B, T* __restrict__ C, int rows, int K, int nnz, int nslots,
                const int* __restrict__ slot_rows) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
  if (w >= nslots) return;
  const int col = (blockIdx.y * 32 + lane) * VEC;
  const bool active = col < K;
  const int lo = w * SPMM_W, hi = min(lo + SPMM_W, nnz);
  const int R0 = __ldg(slot_rows + w), R1 = __ldg(slot_rows + w + 1);

  // (1) the tail of a hub row that started in an earlier slot and covers nonzero `lo`
  if (R0 > 0 && lo < nnz) {
    const int s = __ldg(pos + R0 - 1), e = __ldg(pos + R0);
    if (e > lo && e - s > SPMM_LONG)
      spmm_walk<T, VEC, COLMAJOR, false>(crd, vals, B, C, rows, K, col, active, lane, lo, min(hi, e), R0 - 1, 1, 0);
  }
  // (2) the rows this slot owns, 32 at a time
  for (int rb = R0; rb < R1; rb += 32) {
    const int r = rb + lane;
    const bool valid = r < R1;
    const int s = valid ? __ldg(pos + r) : 0;
    const int e = valid ? __ldg(pos + r + 1) : 0;
    const unsigned hub = __ballot_sync(0xffffffffu, valid && (e - s > SPMM_LONG));
    const int nvalid = min(32, R1 - rb);
    const int nshort = hub ? (__ffs(hub) - 1) : nvalid;   // a hub row is always the last row a slot owns
    if (nshort > 0) {
      const int a = __shfl_sync(0xffffffffu, s, 0);
      const int b = __shfl_sync(0xffffffffu, e, nshort - 1);
      spmm_walk<T, VEC, COLMAJOR, true>(crd, vals, B, C, rows, K, col, active, lane, a, b, rb, nshort, e);
    }
    if (hub) {
      const int h = __ffs(hub) - 1;
      const int hs = __shfl_sync(0xffffffffu, s, h), he = __shfl_sync(0xffffffffu, e, h);
      spmm_walk<T, VEC, COLMAJOR, false>(crd, vals, B, C, rows, K, col, active, lane, hs, min(hi, he), rb + h, 1, 0);
    }
  }
}

template <typename T, int VEC, bool COLMAJOR>
static int spmm_launch_impl(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, int nnz) {
  int nslots = nnz > 0 ? (nnz + SPMM_W - 1) / SPMM_W : 1;
  void* slot_rows = nullptr;
  TB_TRY(scratch_alloc(&slot_rows, sizeof(int) * (size_t)(nslots + 1)));
  spmm_slot_rows_kernel<T, COLMAJOR><<<(nslots + 1 + 255) / 256, 256, 0, stream()>>>(pos, rows, nnz, nslots, K,
                                                                                     (int*)slot_rows, C);
  dim3 grid((nslots + SPMM_WARPS - 1) / SPMM_WARPS, (K + 32 * VEC - 1) / (32 * VEC));
  {
    ProfScope ps("spmm_csr");
    spmm_csr_kernel<T, VEC, COLMAJOR><<<grid, SPMM_WARPS * 32, 0, stream()>>>(pos, crd, vals, B, C, rows, K, nnz, nslots,
                                                                              (const int*)slot_rows);
  }
  count_launch(2);
  scratch_free(slot_rows);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

template <typename T>
static int spmm_launch(const int* pos, const int* crd, const T* vals, const T* B, T* C, int rows, int K, int nnz,
                       bool colmajor) {
  constexpr int V = 16 / sizeof(T);
  bool vec_ok = (K % V == 0) && (((uintptr_t)B & 15) == 0) && (((uintptr_t)C & 15) == 0);
  if (colmajor) {
    if (vec_ok) return spmm_launch_impl<T, V, true>(pos, crd, vals, B, C, rows, K, nnz);
    return spmm_launch_impl<T, 1, true>(pos, crd, vals, B, C, rows, K, nnz);
  }
  if (vec_ok) return spmm_launch_impl<T, V, false>(pos, crd, vals, B, C, rows, K, nnz);
  return spmm_launch_impl<T, 1, false>(pos, crd, vals, B, C, rows, K, nnz);
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
  In pos, crd, vals, bin; Out cout;
  TB_TRY(pos.acquire(Av.pos, sizeof(int32_t) * ((size_t)Av.rows + 1)));
  TB_TRY(crd.acquire(Av.crd ? (void*)Av.crd : (void*)Av.pos, sizeof(int32_t) * (size_t)nnz));
  TB_TRY(vals.acquire(Av.vals ? Av.vals : (void*)Av.pos, es * (size_t)nnz));
  TB_TRY(bin.acquire(Bv.vals, es * (size_t)Av.cols * K));
  TB_TRY(cout.acquire(Cv.vals, es * (size_t)Av.rows * K));
  if (Av.rows > 0 && K > 0) {
    if (Av.dt == DType::F32)
      TB_TRY(spmm_launch<float>(pos.as<int>(), crd.as<int>(), vals.as<float>(), bin.as<float>(), cout.as<float>(),
                                Av.rows, K, nnz, cm));
    else
      TB_TRY(spmm_launch<double>(pos.as<int>(), crd.as<int>(), vals.as<double>(), bin.as<double>(), cout.as<double>(),
                                 Av.rows, K, nnz, cm));
  }
  TB_TRY(cout.commit());
  return finish_call();
}

int taco_b200_spmm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  TB_TRY(taco_b200_spmm_assemble(C, A, B));
  return taco_b200_spmm_compute(C, A, B);
}

}  // extern "C"
