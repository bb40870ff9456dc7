// tools/ref_cuda_bench.cu -- the "recompiled reference kernel" comparator: runs the CUDA source the
// REFERENCE generates for its own GPU schedule (emitted by `oracle/_ref/taco_ref_harness emit_cuda`, compiled for sm_100a by
// oracle/Makefile) on the same operands bench.py uses, the way the reference runs it: every array, and the taco_tensor_t
// structs themselves, in cudaMallocManaged memory.  Test infrastructure: the generated text lives only under oracle/_ref/gen.
//   ref_cuda_spmv <A.tbin>   y(i) = A(i,j) * x(j)      fp64   (scheduleSpMVGPU)
//   ref_cuda_spmm <A.tbin>   C(i,k) = A(i,j) * B(j,k)  fp32, K = 128 (scheduleSpMMGPU)
//   ref_cuda_mttkrp <B.tbin> A(i,j) = B(i,k,l) * C(k,j) * D(l,j)  fp64, rank 32 (scheduleMTTKRPGPU)
// Prints whole-call wall times of compute(); kernel-only times come from running this binary under
// `ncu --metrics gpu__time_duration.sum`.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include REF_SRC            // the reference-generated .cu: taco_tensor_t, prelude kernels, computeDeviceKernel0, compute()
#include "../oracle/tbin.h"

#ifdef REF_SPMM
typedef float val_t;
#else
typedef double val_t;
#endif
// (REF_MTTKRP is fp64 as well)

template <typename T>
static T* managed(size_t n) {
  T* p = nullptr;
  if (cudaMallocManaged((void**)&p, sizeof(T) * (n ? n : 1)) != cudaSuccess) { fprintf(stderr, "cudaMallocManaged failed\n"); exit(1); }
  return p;
}

static taco_tensor_t* make_tensor(int order, const int* dims, const taco_mode_t* types) {
  taco_tensor_t* t = managed<taco_tensor_t>(1);
  t->order = order;
  t->dimensions = managed<int32_t>(order);
  t->mode_ordering = managed<int32_t>(order);
  t->mode_types = managed<taco_mode_t>(order);
  t->indices = managed<uint8_t**>(order);
  for (int l = 0; l < order; l++) {
    t->dimensions[l] = dims[l]; t->mode_ordering[l] = l; t->mode_types[l] = types[l];
    t->indices[l] = managed<uint8_t*>(2);
    t->indices[l][0] = t->indices[l][1] = nullptr;
    if (types[l] == taco_mode_dense) { int32_t* d = managed<int32_t>(1); d[0] = dims[l]; t->indices[l][0] = (uint8_t*)d; }
  }
  t->csize = sizeof(val_t) * 8; t->vals = nullptr; t->fill_value = nullptr; t->vals_size = 0;
  return t;
}

template <typename T>
static T* managed_copy(tbin_array* a) {
  T* p = managed<T>(a->count);
  memcpy(p, a->data, sizeof(T) * a->count);
  free(a->data); a->data = nullptr;
  return p;
}

#if defined(REF_MTTKRP) || defined(REF_TTV) || defined(REF_TTM)
int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <B.tbin> [reps]\n", argv[0]); return 2; }
  const int reps = argc > 2 ? atoi(argv[2]) : 2;
  tbin_file f;
  if (tbin_read(argv[1], &f) != 0) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const int* dims = (const int*)tbin_get(&f, "dims")->data;          // I K L R
  const int I = dims[0], K = dims[1], L = dims[2], R = 32;
  const taco_mode_t sss[3] = {taco_mode_sparse, taco_mode_sparse, taco_mode_sparse}, dd[2] = {taco_mode_dense, taco_mode_dense};
  const int bd[3] = {I, K, L};
  taco_tensor_t* B = make_tensor(3, bd, sss);
#if defined(REF_MTTKRP)
  const int cd[2] = {K, R}, ddm[2] = {L, R}, ad[2] = {I, R};
  taco_tensor_t *C = make_tensor(2, cd, dd), *D = make_tensor(2, ddm, dd), *A = make_tensor(2, ad, dd);
#elif defined(REF_TTV)
  const taco_mode_t d1[1] = {taco_mode_dense};
  const int cd[1] = {L}, ad[2] = {I, K};
  taco_tensor_t *C = make_tensor(1, cd, d1), *A = make_tensor(2, ad, dd);
#else
  const taco_mode_t ddd[3] = {taco_mode_dense, taco_mode_dense, taco_mode_dense};
  const int cd[2] = {L, R}, ad[3] = {I, K, R};
  taco_tensor_t *C = make_tensor(2, cd, dd), *A = make_tensor(3, ad, ddd);
#endif
  const char* pn[3] = {"B1_pos", "B2_pos", "B3_pos"};
  const char* cn[3] = {"B1_crd", "B2_crd", "B3_crd"};
  for (int l = 0; l < 3; l++) {
    B->indices[l][0] = (uint8_t*)managed_copy<int32_t>(tbin_get(&f, pn[l]));
    B->indices[l][1] = (uint8_t*)managed_copy<int32_t>(tbin_get(&f, cn[l]));
  }
  const size_t nnz = tbin_get(&f, "B_vals")->count;
  B->vals = (uint8_t*)managed_copy<double>(tbin_get(&f, "B_vals"));
#if defined(REF_MTTKRP)
  double *c = managed<double>((size_t)K * R), *d = managed<double>((size_t)L * R);
  for (size_t q = 0; q < (size_t)K * R; q++) c[q] = (double)((q * 2654435761u >> 22) & 1023) / 1024;
  for (size_t q = 0; q < (size_t)L * R; q++) d[q] = (double)((q * 40503u >> 6) & 1023) / 1024;
  C->vals = (uint8_t*)c; D->vals = (uint8_t*)d;
  A->vals = (uint8_t*)managed<double>((size_t)I * R);
  const double flops = 3.0 * nnz * R;
#elif defined(REF_TTV)
  double* c = managed<double>((size_t)L);
  for (size_t q = 0; q < (size_t)L; q++) c[q] = (double)((q * 2654435761u >> 22) & 1023) / 1024;
  C->vals = (uint8_t*)c;
  A->vals = (uint8_t*)managed<double>((size_t)I * K);
  const double flops = 2.0 * nnz;
#else
  double* c = managed<double>((size_t)L * R);
  for (size_t q = 0; q < (size_t)L * R; q++) c[q] = (double)((q * 2654435761u >> 22) & 1023) / 1024;
  C->vals = (uint8_t*)c;
  A->vals = (uint8_t*)managed<double>((size_t)I * K * R);
  const double flops = 2.0 * nnz * R;
#endif
  for (int r = 0; r < reps + 1; r++) {
    auto t0 = std::chrono::steady_clock::now();
#if defined(REF_MTTKRP)
    compute(A, B, C, D);
#else
    compute(A, B, C);
#endif
    cudaDeviceSynchronize();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"call\": %d, \"compute_wall_ms\": %.3f, \"gflops\": %.2f}\n", r, ms, flops / ms / 1e6);
    fflush(stdout);
  }
  double s = 0; for (int q = 0; q < 1000; q++) s += ((double*)A->vals)[(size_t)q * R]; printf("checksum %.6g\n", s);
  return 0;
}
#else
int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <A.tbin> [reps]\n", argv[0]); return 2; }
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  tbin_file f;
  if (tbin_read(argv[1], &f) != 0) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
  const int* dims = (const int*)tbin_get(&f, "dims")->data;
  tbin_array *pos = tbin_get(&f, "A_pos"), *crd = tbin_get(&f, "A_crd"), *vals = tbin_get(&f, "A_vals");
  const int n = dims[0], m = dims[1];
  const size_t nnz = crd->count;
  const taco_mode_t csr[2] = {taco_mode_dense, taco_mode_sparse}, dd[2] = {taco_mode_dense, taco_mode_dense}, d1[1] = {taco_mode_dense};
  const int adims[2] = {n, m};
  taco_tensor_t* A = make_tensor(2, adims, csr);
  int32_t* mp = managed<int32_t>(pos->count); memcpy(mp, pos->data, sizeof(int32_t) * pos->count);
  int32_t* mc = managed<int32_t>(nnz); memcpy(mc, crd->data, sizeof(int32_t) * nnz);
  val_t* mv = managed<val_t>(nnz); memcpy(mv, vals->data, sizeof(val_t) * nnz);
  A->indices[1][0] = (uint8_t*)mp; A->indices[1][1] = (uint8_t*)mc; A->vals = (uint8_t*)mv;
#ifdef REF_SPMM
  const int K = 128;
  const int bdims[2] = {m, K}, cdims[2] = {n, K};
  taco_tensor_t *B = make_tensor(2, bdims, dd), *C = make_tensor(2, cdims, dd);
  val_t* b = managed<val_t>((size_t)m * K);
  for (size_t q = 0; q < (size_t)m * K; q++) b[q] = (val_t)((q * 2654435761u >> 22) & 1023) / 1024;
  B->vals = (uint8_t*)b;
  C->vals = (uint8_t*)managed<val_t>((size_t)n * K);
  const double flops = 2.0 * nnz * K;
#else
  const int xdims[1] = {m}, ydims[1] = {n};
  taco_tensor_t *x = make_tensor(1, xdims, d1), *y = make_tensor(1, ydims, d1);
  val_t* xv = managed<val_t>(m);
  for (int q = 0; q < m; q++) xv[q] = (val_t)(((unsigned)q * 2654435761u >> 22) & 1023) / 1024;
  x->vals = (uint8_t*)xv;
  y->vals = (uint8_t*)managed<val_t>(n);
  const double flops = 2.0 * nnz;
#endif
  for (int r = 0; r < reps + 1; r++) {         // call 0 = warm-up (first-touch migration of every operand)
    auto t0 = std::chrono::steady_clock::now();
#ifdef REF_SPMM
    compute(C, A, B);
#else
    compute(y, A, x);
#endif
    cudaDeviceSynchronize();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"call\": %d, \"compute_wall_ms\": %.3f, \"gflops\": %.2f}\n", r, ms, flops / ms / 1e6);
  }
#ifdef REF_SPMM
  double s = 0; for (int q = 0; q < 1000; q++) s += ((val_t*)C->vals)[(size_t)q * 128]; printf("checksum %.6g\n", s);
#else
  double s = 0; for (int q = 0; q < 1000; q++) s += ((val_t*)y->vals)[q]; printf("checksum %.6g\n", s);
#endif
  return 0;
}
#endif
