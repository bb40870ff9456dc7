// spmv.cu -- y(i) = A(i,j) * x(j), A CSR, fp32 / fp64.
//
// Replaces the CUDA the reference emits for SpMV (schedules scheduleSpMVGPU /
// scheduleSpMVSplitPosGPU / scheduleSpMVRowsGPU, /root/reference/test/tests-scheduling-eval.cpp:193-247):
//   reference: nnz-split over fused pos space, per-thread binary search, ONE GLOBAL fp64 atomicAdd PER NONZERO,
//              host-serial zeroing of y, block-start array allocated+freed per call in managed memory.
//   here     : nnz-balanced ROW-ALIGNED tiles (tile b owns the rows whose first nonzero falls in
//              [b*TILE, (b+1)*TILE) -- same binary search as taco_binarySearchBeforeBlock,
//              /root/reference/src/codegen/codegen_cuda.cpp:110-125, but each row has exactly one owner so
//              no atomics and no zero-fill pass are needed); crd/vals are streamed with aligned 128-bit
//              ld.global.nc.L1::no_allocate loads, products are staged in shared memory and every row is
//              reduced by one thread in ascending position order -- the operation order of the reference's C
//              kernel (Appendix A.1), so results are bit-identical to the oracle, not merely within 1e-12.
// Algorithmic bytes per launch: nnz*(4+sizeof T) + 4(n+1) + sizeof T*(cols + rows).
#include "common.cuh"

namespace tb {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_VEC = 4;                 // nonzeros per thread per vector step (one int4 of crd)

// tile_rows[b] = first row r with pos[r] >= b*tile   (b = 0..ntiles-1), tile_rows[ntiles] = rows
__global__ void spmv_tile_rows_kernel(const int* __restrict__ pos, int rows, int tile, int ntiles,
                                      int* __restrict__ tile_rows) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > ntiles) return;
  tile_rows[b] = (b == ntiles) ? rows : tbd::search_first_ge(pos, 0, rows, b * tile);
}

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

// STEPS vector steps per thread => tile of THREADS*VEC*STEPS nonzeros staged in shared memory.
template <typename T, int STEPS>
__global__ void __launch_bounds__(SPMV_THREADS)
spmv_csr_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                const T* __restrict__ x, T* __restrict__ y, const int* __restrict__ tile_rows, int nnz) {
  constexpr int TILE = SPMV_THREADS * SPMV_VEC * STEPS;
  __shared__ T prod[TILE + SPMV_VEC];
  const int r0 = tile_rows[blockIdx.x], r1 = tile_rows[blockIdx.x + 1];
  if (r0 >= r1) return;
  const int p0 = __ldg(pos + r0), p1 = __ldg(pos + r1);
  const int tid = threadIdx.x;

  if (p1 - (p0 & ~(SPMV_VEC - 1)) <= TILE) {
    // ---- fast path: the whole tile fits one staging pass ------------------------------------------------
    const int base = p0 & ~(SPMV_VEC - 1);          // 16-byte aligned start for the vector loads
    int4 c[STEPS];
    T v[STEPS][4];
#pragma unroll
    for (int s = 0; s < STEPS; s++) {
      int q = base + (s * SPMV_THREADS + tid) * SPMV_VEC;
      if (q + SPMV_VEC <= nnz) {
        c[s] = tbd::ldg_stream_i4(crd + q);
        ValVec<T>::load4(vals + q, v[s]);
      } else {
        int cc[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          bool ok = q + e < nnz;
          cc[e] = ok ? __ldg(crd + q + e) : 0;
          v[s][e] = ok ? __ldg(vals + q + e) : T(0);
        }
        c[s] = make_int4(cc[0], cc[1], cc[2], cc[3]);
      }
    }
#pragma unroll
    for (int s = 0; s < STEPS; s++) {
      int q = base + (s * SPMV_THREADS + tid) * SPMV_VEC;
      if (q < p1) {   // entries in [base,p0) and [p1, ..) are loaded but never consumed
        T x0 = __ldg(x + c[s].x), x1 = __ldg(x + c[s].y), x2 = __ldg(x + c[s].z), x3 = __ldg(x + c[s].w);
        T* d = prod + (q - base);
        d[0] = v[s][0] * x0; d[1] = v[s][1] * x1; d[2] = v[s][2] * x2; d[3] = v[s][3] * x3;
      }
    }
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += SPMV_THREADS) {
      int s = __ldg(pos + r) - base, e = __ldg(pos + r + 1) - base;
      T acc = T(0);
      for (int q = s; q < e; q++) acc += prod[q];
      y[r] = acc;
    }
    return;
  }

  // ---- slow path: the tile owns a row longer than the staging buffer; walk it in chunks -----------------
  // Empty rows first (they never intersect a chunk).
  for (int r = r0 + tid; r < r1; r += SPMV_THREADS)
    if (__ldg(pos + r) == __ldg(pos + r + 1)) y[r] = T(0);
  T carry = T(0);
  int carry_row = -1;
  for (int lo = p0; lo < p1; lo += TILE) {
    const int hi = min(lo + TILE, p1);
    __syncthreads();
    for (int q = lo + tid; q < hi; q += SPMV_THREADS) prod[q - lo] = __ldg(vals + q) * __ldg(x + __ldg(crd + q));
    __syncthreads();
    const int rf = tbd::search_last_le(pos, r0, r1 - 1, lo);       // row containing (or preceding) lo
    const int rl = tbd::search_last_le(pos, r0, r1 - 1, hi - 1);   // row containing hi-1
    // rows are owned by (r - r0) % THREADS so a row straddling two chunks stays with the thread holding its carry
    for (int r = rf + ((tid - (rf - r0) % SPMV_THREADS + SPMV_THREADS) % SPMV_THREADS); r <= rl; r += SPMV_THREADS) {
      int rs = __ldg(pos + r), re = __ldg(pos + r + 1);
      if (rs == re) continue;
      int s = max(rs, lo), e = min(re, hi);
      if (s >= e) continue;
      T acc = (carry_row == r) ? carry : T(0);
      for (int q = s; q < e; q++) acc += prod[q - lo];
      if (re <= hi) y[r] = acc;
      else { carry = acc; carry_row = r; }
    }
  }
}

template <typename T>
static int spmv_launch(const CsrView& A, const In& pos, const In& crd, const In& vals, const In& x, Out& y, int nnz) {
  constexpr int STEPS = 2;
  constexpr int TILE = SPMV_THREADS * SPMV_VEC * STEPS;
  int ntiles = nnz > 0 ? (nnz + TILE - 1) / TILE : 1;
  void* tile_rows = nullptr;
  TB_TRY(scratch_alloc(&tile_rows, sizeof(int) * (size_t)(ntiles + 1)));
  spmv_tile_rows_kernel<<<(ntiles + 1 + 255) / 256, 256, 0, stream()>>>(pos.as<int>(), A.rows, TILE, ntiles,
                                                                          (int*)tile_rows);
  {
    ProfScope ps("spmv_csr");
    spmv_csr_kernel<T, STEPS><<<ntiles, SPMV_THREADS, 0, stream()>>>(pos.as<int>(), crd.as<int>(), vals.as<T>(),
                                                                      x.as<T>(), y.as<T>(), (const int*)tile_rows, nnz);
  }
  count_launch(2);
  scratch_free(tile_rows);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

int csr_nnz(const CsrView& A, int32_t vals_size_hint, int32_t* nnz) {
  if (!A.pos) return fail(TACO_B200_ERR_ARG, "CSR operand has no pos array");
  if (classify(A.pos) == Mem::Device && vals_size_hint > 0) { *nnz = vals_size_hint; return TACO_B200_OK; }
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
  if (((uintptr_t)crd.dptr | (uintptr_t)vals.dptr) & 15)
    return fail(TACO_B200_ERR_ARG, "spmv: device crd/vals arrays must be 16-byte aligned");
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
