// csf.cu -- order-3 CSF kernels: MTTKRP, TTV, TTM.  B is {Compressed,Compressed,Compressed}, mode order 0,1,2.
//
//   MTTKRP  A(i,j)   = B(i,k,l) * C(k,j) * D(l,j)      replaces scheduleMTTKRPGPU (tests-scheduling-eval.cpp:327-342)
//   TTV     A(i,j)   = B(i,j,k) * c(k)                 replaces scheduleTTVGPU    (:308-325)
//   TTM     A(i,j,l) = B(i,j,k) * C(k,l)               replaces scheduleTTMGPU    (:289-306)
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

struct CsfCall {
  Csf3View B; DType dt; int32_t nslices, nfib, nnz;
  In p1, c1, p2, c2, p3, c3, vals;
};

static int csf_prepare(taco_tensor_t* Bt, CsfCall* cc) {
  TB_TRY(view_csf3(Bt, "B", &cc->B));
  cc->dt = cc->B.dt;
  // level sizes: pos1[1], pos2[nslices], pos3[nfib]  (host arrays are read directly; device arrays cost one
  // small read-back each -- callers on the hot loop keep B host-described or pass sizes through vals_size)
  int32_t p10 = 0;
  TB_TRY(read_i32(cc->B.pos[0], &p10));
  TB_TRY(read_i32(cc->B.pos[0] + 1, &cc->nslices));
  cc->nslices -= p10;
  TB_TRY(read_i32(cc->B.pos[1] + cc->nslices, &cc->nfib));
  if (classify(cc->B.pos[2]) == Mem::Device && Bt->vals_size > 0) cc->nnz = Bt->vals_size;
  else TB_TRY(read_i32(cc->B.pos[2] + cc->nfib, &cc->nnz));
  if (cc->nslices < 0 || cc->nfib < 0 || cc->nnz < 0) return fail(TACO_B200_ERR_ARG, "csf: corrupt pos arrays");
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

template <typename T>
static int mttkrp_launch(CsfCall& cc, const T* C, const T* D, T* A, size_t a_count, int R) {
  if (cc.nslices == 0 || R == 0) {
    TB_CUDA(cudaMemsetAsync(A, 0, a_count * sizeof(T), stream()));
    count_launch(1);
  }
  if (cc.nslices > 0 && R > 0) {
    long long ctas = ((long long)cc.nslices + CSF_WARPS - 1) / CSF_WARPS;
    int grid = (int)(ctas < (1 << 22) ? ctas : (1 << 22));
    ProfScope ps("mttkrp_csf");
    csf3_rows_kernel<T, 0><<<grid, CSF_WARPS * 32, 0, stream()>>>(cc.p1.as<int>(), cc.c1.as<int>(), cc.p2.as<int>(),
        cc.c2.as<int>(), cc.p3.as<int>(), cc.c3.as<int>(), cc.vals.as<T>(), C, D, A, R, 0, cc.B.dim[0]);
    count_launch(1);
  }
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

template <typename T>
static int ttm_launch(CsfCall& cc, const T* C, T* A, size_t a_count, int R, int Kdim) {
  TB_CUDA(cudaMemsetAsync(A, 0, a_count * sizeof(T), stream()));
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

template <typename T>
static int ttv_launch(CsfCall& cc, const T* c, T* A, size_t a_count, int Kdim) {
  TB_CUDA(cudaMemsetAsync(A, 0, a_count * sizeof(T), stream()));
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
