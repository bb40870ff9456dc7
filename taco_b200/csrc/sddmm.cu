// sddmm.cu -- A(i,j) = B(i,j) * C(i,k) * D(j,k); A, B CSR (A takes B's structure); C, D dense row-major.
//
// Replaces the CUDA the reference emits for scheduleSDDMMGPU (/root/reference/test/tests-scheduling-eval.cpp:270-287):
//   reference: dense n x n output (impossible at config C3), warp walks 256 nnz, lanes split the contraction,
//              atomicAddWarp (shuffle tree + one global atomic) per (nnz, dense_val), binary search per thread.
//   here     : CSR output in B's structure (SURVEY.md Appendix A.1 "SDDMM, CSR output"), pure nnz-split (every
//              output value is independent, so no reduction across slots): a group of G = K/VEC lanes owns one
//              nonzero, reads the C row (L1-resident across the row's nonzeros) and gathers the D row with
//              128-bit loads, reduces inside the group with xor shuffles and the warp writes 32 results with one
//              coalesced store.  Row of each nonzero: one binary search per nonzero inside the slot's row range.
//   assemble : the result structure is B's (append assembly of the reference yields exactly B's pos/crd,
//              src/lower/lowerer_impl_imperative.cpp:3157-3308): two device copies.
// Algorithmic bytes per launch: nnz*(4 + 2*sizeof T) + 4(n+1) + 2*sizeof T*n*K (+ 8*nnz+4(n+1) when assembling).
#include <cstdlib>
#include <cstring>

#include <thread>
#include <vector>

#include "common.cuh"

namespace tb {

constexpr int SDDMM_W = 128;       // nonzeros per warp slot
constexpr int SDDMM_WARPS = 8;

__global__ void slot_first_row_kernel(const int* __restrict__ pos, int rows, int nnz, int slot, int nslots,
                                      int* __restrict__ first) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > nslots) return;
  long long lo = (long long)w * slot;
  first[w] = (w == nslots || lo >= nnz) ? max(rows - 1, 0) : tbd::search_last_le(pos, 0, rows, (int)lo);
}

template <typename T, int VEC> struct Ld;
template <> struct Ld<float, 4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};
template <> struct Ld<double, 2> {
  static __device__ __forceinline__ void ld(const double* p, double (&v)[2]) {
    double2 a = __ldg(reinterpret_cast<const double2*>(p)); v[0] = a.x; v[1] = a.y;
  }
};
template <typename T> struct Ld<T, 1> {
  static __device__ __forceinline__ void ld(const T* p, T (&v)[1]) { v[0] = __ldg(p); }
};

template <typename T, int VEC, int G>
__global__ void __launch_bounds__(SDDMM_WARPS * 32)
sddmm_csr_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ bvals,
                 const T* __restrict__ C, const T* __restrict__ D, T* __restrict__ avals, int rows, int K, int nnz,
                 int nslots, const int* __restrict__ slot_first) {
  constexpr int NG = 32 / G;           // nonzeros processed concurrently by a warp
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * SDDMM_WARPS + (threadIdx.x >> 5);
  if (w >= nslots) return;
  const int lo = w * SDDMM_W, hi = min(lo + SDDMM_W, nnz);
  const int rlo = __ldg(slot_first + w), rhi = __ldg(slot_first + w + 1);
  const int g = lane / G, gl = lane % G;
  for (int base = lo; base < hi; base += 32) {
    const int p = base + lane;
    const bool ok = p < hi;
    int my_row = 0, my_col = 0;
    T my_b = T(0);
    if (ok) {
      my_row = tbd::search_last_le(pos, rlo, rhi, p);
      my_col = tbd::ldg_stream_i32(crd + p);
      my_b = __ldg(bvals + p);
    }
    T res[G];
#pragma unroll
    for (int t = 0; t < G; t++) {
      const int q = t * NG + g;
      const int i = __shfl_sync(0xffffffffu, my_row, q);
      const int j = __shfl_sync(0xffffffffu, my_col, q);
      const T b = __shfl_sync(0xffffffffu, my_b, q);
      const T* c = C + (size_t)i * K;
      const T* d = D + (size_t)j * K;
      T part = T(0);
      for (int kk = gl * VEC; kk < K; kk += G * VEC) {
        T cv[VEC], dv[VEC];
        Ld<T, VEC>::ld(c + kk, cv);
        Ld<T, VEC>::ld(d + kk, dv);
#pragma unroll
        for (int x = 0; x < VEC; x++) part = part + (b * cv[x]) * dv[x];     // reference association (B*C)*D
      }
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
      res[t] = part;
    }
    T out = T(0);
#pragma unroll
    for (int t = 0; t < G; t++) {
      T v = __shfl_sync(0xffffffffu, res[t], (lane % NG) * G);
      if (lane / NG == t) out = v;
    }
    if (ok) avals[p] = out;
  }
}

// ---- single-chunk kernel: K == G*VEC (a dense row is exactly one 16-byte piece per lane of a G-lane group) -----------
// The C row of a group stays in registers while the nonzeros of one row go by, U gathered D rows are in flight per
// group, and the U partial dot products a lane holds are reduced across the group with a transposing butterfly
// (U-1 + log2(G/U) shuffles for U nonzeros instead of U*log2(G)).
template <typename T, int U, int G>
__device__ __forceinline__ T sddmm_multi_reduce(T (&v)[U], int gl) {
  int off = G / 2;
#pragma unroll
  for (int n = U; n > 1; n >>= 1) {
    const int half = n / 2;
    const bool upper = (gl & off) != 0;
#pragma unroll
    for (int t = 0; t < half; t++) {
      const T send = upper ? v[t] : v[t + half];
      const T keep = upper ? v[t + half] : v[t];
      v[t] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];      // lane gl holds the total of nonzero u = gl / (G/U)
}

template <typename T, int VEC>
__device__ __forceinline__ void sddmm_ld_keep(const T* p, T (&v)[VEC], uint64_t keep) {
  if constexpr (VEC == 4 && sizeof(T) == 4)
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p), "l"(keep));
  else if constexpr (VEC == 2 && sizeof(T) == 8)
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[0]), "=d"(v[1]) : "l"(p), "l"(keep));
  else
    v[0] = __ldg(p);
}

template <typename T, int VEC, int G, int UREQ, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
sddmm_csr_chunk_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ bvals,
                       const T* __restrict__ C, const T* __restrict__ D, T* __restrict__ avals, int K, int nnz, int nslots,
                       const int* __restrict__ slot_first) {
  constexpr int U = UREQ < G ? UREQ : G;     // nonzeros in flight per lane group
  constexpr int ROUNDS = G / U;              // rounds per batch of 32 nonzeros
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (w >= nslots) return;
  const uint64_t keep = tbd::policy_evict_last(), strm = tbd::policy_evict_first();
  const int lo = w * SDDMM_W, hi = min(lo + SDDMM_W, nnz);
  const int rlo = __ldg(slot_first + w), rhi = __ldg(slot_first + w + 1);
  const int g = lane / G, gl = lane % G;
  const T* Cg = C + gl * VEC;
  const T* Dg = D + gl * VEC;
  T c[VEC];
#pragma unroll
  for (int x = 0; x < VEC; x++) c[x] = T(0);
  int cur_i = -1;
  for (int base = lo; base < hi; base += 32) {
    const int p = base + lane;
    const bool ok = p < hi;
    int my_row = rlo, my_col = 0;
    T my_b = T(0);
    if (ok) {
      my_row = tbd::search_last_le(pos, rlo, rhi, p);
      my_col = tbd::ldg_stream_i32(crd + p, strm);
      my_b = tbd::ldg_stream(bvals + p, strm);
    }
    T res[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      T d[U][VEC];
      int ii[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int q = g * G + r * U + u;
        const int j = __shfl_sync(0xffffffffu, my_col, q);
        ii[u] = __shfl_sync(0xffffffffu, my_row, q);
        sddmm_ld_keep<T, VEC>(Dg + (size_t)j * K, d[u], keep);
      }
      T part[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int q = g * G + r * U + u;
        if (ii[u] != cur_i) {              // next row of the sparse operand: refresh the group's piece of C(i,:)
          cur_i = ii[u];
          sddmm_ld_keep<T, VEC>(Cg + (size_t)cur_i * K, c, keep);
        }
        const T b = __shfl_sync(0xffffffffu, my_b, q);
        T acc = T(0);
#pragma unroll
        for (int x = 0; x < VEC; x++) acc = acc + (b * c[x]) * d[u][x];     // reference association (B*C)*D
        part[u] = acc;
      }
      res[r] = sddmm_multi_reduce<T, U, G>(part, gl);
    }
    // nonzero (round r, slot u) of group g now sits in lanes g*G + u*ROUNDS .. +ROUNDS-1; lane u*ROUNDS+r forwards round r
    T send = res[0];
#pragma unroll
    for (int r = 1; r < ROUNDS; r++)
      if (gl % ROUNDS == r) send = res[r];
    const T out = __shfl_sync(0xffffffffu, send, g * G + (gl % U) * ROUNDS + gl / U);
    if (ok) {
      if constexpr (sizeof(T) == 8) asm volatile("st.global.L1::no_allocate.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(avals + p), "d"(out), "l"(strm) : "memory");
      else asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(avals + p), "f"(out), "l"(strm) : "memory");
    }
  }
}

template <typename T, int VEC, int G>
static void sddmm_chunk_go(int variant, int grid8, int grid4, const int* pos, const int* crd, const T* bvals, const T* C, const T* D,
                           T* avals, int K, int nnz, int nslots, const int* first) {
  switch (variant) {
    case 1: sddmm_csr_chunk_kernel<T, VEC, G, 2, 8, 6><<<grid8, 256, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
    case 2: sddmm_csr_chunk_kernel<T, VEC, G, 8, 8, 3><<<grid8, 256, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
    case 3: sddmm_csr_chunk_kernel<T, VEC, G, 16, 8, 2><<<grid8, 256, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
    case 4: sddmm_csr_chunk_kernel<T, VEC, G, 16, 8, 3><<<grid8, 256, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
    case 5: sddmm_csr_chunk_kernel<T, VEC, G, 8, 4, 8><<<grid4, 128, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
    case 6: sddmm_csr_chunk_kernel<T, VEC, G, 8, 8, 5><<<grid8, 256, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
    default: sddmm_csr_chunk_kernel<T, VEC, G, 8, 8, 4><<<grid8, 256, 0, stream()>>>(pos, crd, bvals, C, D, avals, K, nnz, nslots, first); break;
  }
}

template <typename T, int VEC>
static int sddmm_launch_g(const int* pos, const int* crd, const T* bvals, const T* C, const T* D, T* avals, int rows,
                          int K, int nnz) {
  if (nnz == 0) return TACO_B200_OK;
  int nslots = (nnz + SDDMM_W - 1) / SDDMM_W;
  void* first = nullptr;
  TB_TRY(scratch_alloc(&first, sizeof(int) * (size_t)(nslots + 1)));
  slot_first_row_kernel<<<(nslots + 1 + 255) / 256, 256, 0, stream()>>>(pos, rows, nnz, SDDMM_W, nslots, (int*)first);
  int groups = (K + VEC - 1) / VEC;
  int grid = (nslots + SDDMM_WARPS - 1) / SDDMM_WARPS;
#define TB_SDDMM_GO(GG)                                                                                              \
  sddmm_csr_kernel<T, VEC, GG><<<grid, SDDMM_WARPS * 32, 0, stream()>>>(pos, crd, bvals, C, D, avals, rows, K, nnz, \
                                                                        nslots, (const int*)first)
  static const int variant = getenv("TACO_B200_SDDMM_VARIANT") ? atoi(getenv("TACO_B200_SDDMM_VARIANT")) : 0;
  const bool chunk = variant >= 0 && K % VEC == 0 && groups <= 32 && (groups & (groups - 1)) == 0;
  if (chunk) {
    ProfScope ps("sddmm_csr");
    const int g4 = (nslots + 3) / 4;
#define TB_SDDMM_CHUNK(GG) sddmm_chunk_go<T, VEC, GG>(variant, grid, g4, pos, crd, bvals, C, D, avals, K, nnz, nslots, (const int*)first)
    switch (groups) {
      case 1: TB_SDDMM_CHUNK(1); break;
      case 2: TB_SDDMM_CHUNK(2); break;
      case 4: TB_SDDMM_CHUNK(4); break;
      case 8: TB_SDDMM_CHUNK(8); break;
      case 16: TB_SDDMM_CHUNK(16); break;
      default: TB_SDDMM_CHUNK(32); break;
    }
#undef TB_SDDMM_CHUNK
  } else {
  ProfScope ps("sddmm_csr");
  if (groups <= 1) TB_SDDMM_GO(1);
  else if (groups <= 2) TB_SDDMM_GO(2);
  else if (groups <= 4) TB_SDDMM_GO(4);
  else if (groups <= 8) TB_SDDMM_GO(8);
  else if (groups <= 16) TB_SDDMM_GO(16);
  else TB_SDDMM_GO(32);
  }
#undef TB_SDDMM_GO
  count_launch(2);
  scratch_free(first);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

template <typename T>
static int sddmm_launch(const int* pos, const int* crd, const T* bvals, const T* C, const T* D, T* avals, int rows,
                        int K, int nnz) {
  constexpr int V = 16 / sizeof(T);
  bool vec_ok = (K % V == 0) && (((uintptr_t)C | (uintptr_t)D) & 15) == 0;
  if (vec_ok) return sddmm_launch_g<T, V>(pos, crd, bvals, C, D, avals, rows, K, nnz);
  return sddmm_launch_g<T, 1>(pos, crd, bvals, C, D, avals, rows, K, nnz);
}

int csr_nnz(const CsrView& A, int32_t vals_size_hint, int32_t* nnz);   // spmv.cu

struct SddmmViews { CsrView A, B; DenseView C, D; };

static int sddmm_views(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D, SddmmViews* v) {
  TB_TRY(ensure_init());
  TB_TRY(view_csr(A, "A", &v->A));
  TB_TRY(view_csr(B, "B", &v->B));
  TB_TRY(view_dense(C, 2, "C", &v->C));
  TB_TRY(view_dense(D, 2, "D", &v->D));
  if (v->C.mode_order[0] != 0 || v->D.mode_order[0] != 0)
    return fail(TACO_B200_ERR_FORMAT, "sddmm: C and D must be row-major {Dense,Dense}");
  if (v->A.rows != v->B.rows || v->A.cols != v->B.cols || v->C.dim[0] != v->B.rows || v->D.dim[0] != v->B.cols ||
      v->C.dim[1] != v->D.dim[1])
    return fail(TACO_B200_ERR_ARG, "sddmm: dimension mismatch");
  if (v->A.dt != v->B.dt || v->C.dt != v->B.dt || v->D.dt != v->B.dt)
    return fail(TACO_B200_ERR_FORMAT, "sddmm: mixed component types");
  return TACO_B200_OK;
}

// copy `bytes` from an operand array (host or device) into a freshly allocated result array
static int clone_to_result(const void* src, size_t bytes, void** out) {
  void* dst = result_alloc(bytes);
  if (!dst) return fail(TACO_B200_ERR_ALLOC, "cannot allocate %zu result bytes", bytes);
  *out = dst;
  if (bytes == 0) return TACO_B200_OK;
  bool sdev = classify(src) == Mem::Device, ddev = result_space() == TACO_B200_SPACE_DEVICE;
  if (!sdev && !ddev) {                          // host -> fresh host array: parallel copy (takes the page faults in parallel)
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = bytes < (8u << 20) ? 1 : (hw >= 8 ? 8 : (hw > 1 ? (int)hw : 1));
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) {
      const size_t a = (size_t)t * per;
      if (a >= bytes) break;
      const size_t n = bytes - a < per ? bytes - a : per;
      th.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, n); });
    }
    memcpy(dst, src, bytes < per ? bytes : per);
    for (auto& x : th) x.join();
    return TACO_B200_OK;
  }
  if (sdev && !ddev) return d2h_fresh(dst, src, bytes);
  TB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream()));
  return TACO_B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Dense-result form, the statement of the reference's sddmmGPU test (test/tests-scheduling-eval.cpp:1360-1418, schedule
// scheduleSDDMMGPU :270-287):  A(i,k) = B(i,k) * C(i,j) * D(j,k)  with A {Dense,Dense} and D stored (contraction, column).
// D is transposed once on the device (tiled through shared memory) so that the sampled dot products read contiguous rows
// and the tuned CSR kernel runs on it; its values are then scattered into the zero-filled dense result.
// Reference order per entry: j ascending, (B*C)*D -- the kernel's lane-parallel dot product stays within the tolerance.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ in, T* __restrict__ out, int rows, int cols) {
  __shared__ T tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8 threads
  for (int r = ty; r < 32; r += 8)
    if (by + r < rows && bx + tx < cols) tile[r][tx] = __ldg(in + (size_t)(by + r) * cols + bx + tx);
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bx + r < cols && by + tx < rows) out[(size_t)(bx + r) * rows + by + tx] = tile[tx][r];
}

template <typename T>
__global__ void __launch_bounds__(256) scatter_rows_kernel(const int* __restrict__ pos, const int* __restrict__ crd,
                                                           const T* __restrict__ vals, T* __restrict__ A, int rows, int cols) {
  const int warp = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int p1 = __ldg(pos + warp + 1);
  for (int p = __ldg(pos + warp) + lane; p < p1; p += 32) A[(size_t)warp * cols + __ldg(crd + p)] = vals[p];
}

template <typename T>
static int sddmm_dense_run(const int* pos, const int* crd, const T* bvals, const T* C, const T* D, T* A, int rows, int cols, int J,
                           int nnz) {
  TB_CUDA(cudaMemsetAsync(A, 0, sizeof(T) * (size_t)rows * cols, stream()));
  count_launch(1);
  if (nnz == 0 || rows == 0) return TACO_B200_OK;
  void *Dt = nullptr, *vals = nullptr;
  TB_TRY(scratch_alloc(&Dt, sizeof(T) * (size_t)J * cols));
  TB_TRY(scratch_alloc(&vals, sizeof(T) * (size_t)nnz));
  if (J > 0) {
    dim3 grid((cols + 31) / 32, (J + 31) / 32);
    transpose_kernel<T><<<grid, 256, 0, stream()>>>(D, (T*)Dt, J, cols);       // D (J x cols) -> Dt (cols x J)
    count_launch(1);
  }
  int rc = sddmm_launch<T>(pos, crd, bvals, C, (const T*)Dt, (T*)vals, rows, J, nnz);
  if (rc == TACO_B200_OK) {
    const size_t threads = (size_t)rows * 32;
    scatter_rows_kernel<T><<<(unsigned)((threads + 255) / 256), 256, 0, stream()>>>(pos, crd, (const T*)vals, A, rows, cols);
    count_launch(1);
  }
  scratch_free(Dt);
  scratch_free(vals);
  TB_TRY(rc);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

struct SddmmDenseViews { DenseView A, C, D; CsrView B; };
static int sddmm_dense_views(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D, SddmmDenseViews* v) {
  TB_TRY(ensure_init());
  TB_TRY(view_dense(A, 2, "A", &v->A));
  TB_TRY(view_csr(B, "B", &v->B));
  TB_TRY(view_dense(C, 2, "C", &v->C));
  TB_TRY(view_dense(D, 2, "D", &v->D));
  if (v->A.mode_order[0] != 0 || v->C.mode_order[0] != 0 || v->D.mode_order[0] != 0)
    return fail(TACO_B200_ERR_FORMAT, "sddmm_dense: A, C and D must be row-major {Dense,Dense}");
  if (v->A.dim[0] != v->B.rows || v->A.dim[1] != v->B.cols || v->C.dim[0] != v->B.rows || v->D.dim[1] != v->B.cols ||
      v->C.dim[1] != v->D.dim[0])
    return fail(TACO_B200_ERR_ARG, "sddmm_dense: dimension mismatch A[%d x %d] = B[%d x %d] * C[%d x %d] * D[%d x %d]", v->A.dim[0],
                v->A.dim[1], v->B.rows, v->B.cols, v->C.dim[0], v->C.dim[1], v->D.dim[0], v->D.dim[1]);
  if (v->A.dt != v->B.dt || v->C.dt != v->B.dt || v->D.dt != v->B.dt)
    return fail(TACO_B200_ERR_FORMAT, "sddmm_dense: mixed component types");
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_sddmm_dense_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  SddmmDenseViews v;
  TB_TRY(sddmm_dense_views(A, B, C, D, &v));
  void* p = result_alloc(v.A.count() * dsize(v.A.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "sddmm_dense: cannot allocate result");
  A->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

int taco_b200_sddmm_dense_compute(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  SddmmDenseViews v;
  TB_TRY(sddmm_dense_views(A, B, C, D, &v));
  if (!v.A.vals) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
  int32_t nnz = 0;
  TB_TRY(csr_nnz(v.B, B->vals_size, &nnz));
  if (nnz < 0) return fail(TACO_B200_ERR_ARG, "sddmm_dense: negative nnz");
  const int J = v.C.dim[1];
  const size_t es = dsize(v.B.dt);
  In pos, crd, bvals, cin, din; Out aout;
  TB_TRY(pos.acquire(v.B.pos, sizeof(int32_t) * ((size_t)v.B.rows + 1)));
  TB_TRY(crd.acquire(v.B.crd ? (void*)v.B.crd : (void*)v.B.pos, sizeof(int32_t) * (size_t)nnz));
  TB_TRY(bvals.acquire(v.B.vals ? v.B.vals : (void*)v.B.pos, es * (size_t)nnz));
  TB_TRY(cin.acquire(v.C.vals, es * (size_t)v.B.rows * J));
  TB_TRY(din.acquire(v.D.vals, es * (size_t)J * v.B.cols));
  TB_TRY(aout.acquire(v.A.vals, es * v.A.count()));
  ProfScope ps("sddmm_dense");
  if (v.B.dt == DType::F32)
    TB_TRY(sddmm_dense_run<float>(pos.as<int>(), crd.as<int>(), bvals.as<float>(), cin.as<float>(), din.as<float>(), aout.as<float>(),
                                  v.B.rows, v.B.cols, J, nnz));
  else
    TB_TRY(sddmm_dense_run<double>(pos.as<int>(), crd.as<int>(), bvals.as<double>(), cin.as<double>(), din.as<double>(),
                                   aout.as<double>(), v.B.rows, v.B.cols, J, nnz));
  TB_TRY(aout.commit());
  return finish_call();
}

int taco_b200_sddmm_dense_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  TB_TRY(taco_b200_sddmm_dense_assemble(A, B, C, D));
  return taco_b200_sddmm_dense_compute(A, B, C, D);
}

int taco_b200_sddmm_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  SddmmViews v;
  TB_TRY(sddmm_views(A, B, C, D, &v));
  int32_t nnz = 0;
  TB_TRY(csr_nnz(v.B, B->vals_size, &nnz));
  void *pos = nullptr, *crd = nullptr;
  TB_TRY(clone_to_result(v.B.pos, sizeof(int32_t) * ((size_t)v.B.rows + 1), &pos));
  TB_TRY(clone_to_result(v.B.crd ? (void*)v.B.crd : (void*)v.B.pos, sizeof(int32_t) * (size_t)nnz, &crd));
  void* vals = result_alloc(dsize(v.B.dt) * (size_t)nnz);
  if (!vals) return fail(TACO_B200_ERR_ALLOC, "sddmm: cannot allocate result values");
  A->indices[1][0] = (uint8_t*)pos;
  A->indices[1][1] = (uint8_t*)crd;
  A->vals = (uint8_t*)vals;
  A->vals_size = nnz;
  return TACO_B200_OK;
}

int taco_b200_sddmm_compute(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  SddmmViews v;
  TB_TRY(sddmm_views(A, B, C, D, &v));
  int32_t nnz = 0;
  TB_TRY(csr_nnz(v.B, B->vals_size, &nnz));
  if (nnz < 0) return fail(TACO_B200_ERR_ARG, "sddmm: negative nnz");
  const int K = v.C.dim[1];
  size_t es = dsize(v.B.dt);
  In pos, crd, bvals, cin, din; Out aout;
  TB_TRY(pos.acquire(v.B.pos, sizeof(int32_t) * ((size_t)v.B.rows + 1)));
  TB_TRY(crd.acquire(v.B.crd ? (void*)v.B.crd : (void*)v.B.pos, sizeof(int32_t) * (size_t)nnz));
  TB_TRY(bvals.acquire(v.B.vals ? v.B.vals : (void*)v.B.pos, es * (size_t)nnz));
  TB_TRY(cin.acquire(v.C.vals, es * (size_t)v.B.rows * K));
  TB_TRY(din.acquire(v.D.vals, es * (size_t)v.B.cols * K));
  if (nnz > 0) {
    TB_TRY(aout.acquire(v.A.vals, es * (size_t)nnz));
    if (v.B.dt == DType::F32)
      TB_TRY(sddmm_launch<float>(pos.as<int>(), crd.as<int>(), bvals.as<float>(), cin.as<float>(), din.as<float>(),
                                 aout.as<float>(), v.B.rows, K, nnz));
    else
      TB_TRY(sddmm_launch<double>(pos.as<int>(), crd.as<int>(), bvals.as<double>(), cin.as<double>(), din.as<double>(),
                                  aout.as<double>(), v.B.rows, K, nnz));
    TB_TRY(aout.commit());
  }
  return finish_call();
}

int taco_b200_sddmm_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D) {
  TB_TRY(taco_b200_sddmm_assemble(A, B, C, D));
  return taco_b200_sddmm_compute(A, B, C, D);
}

}  // extern "C"
