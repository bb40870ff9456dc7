// tools/ref_cuda_bench.cu -- the "recompiled reference kernel" comparator: runs the CUDA source the
// REFERENCE generates for its own GPU schedule (emitted by `oracle/_ref/taco_ref_harness emit_cuda`, compiled for sm_100a by
// oracle/Makefile) on the same operands bench.py uses, the way the reference runs it: every array, and the taco_tensor_t
// structs themselves, in cudaMallocManaged memory.  Test infrastructure: the generated text lives only under oracle/_ref/gen.
//   ref_cuda_spmv <A.tbin>   y(i) = A(i,j) * x(j)      fp64   (scheduleSpMVGPU)
//   ref_cuda_spmm <A.tbin>   C(i,k) = A(i,j) * B(j,k)  fp32, K = 128 (scheduleSpMMGPU)
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
