/* oracle/taco_oracle.c -- TEST INFRASTRUCTURE ONLY.  See taco_oracle.h for the role and the parity pin.
 *
 * Every loop nest below restates the C that the reference emits for the statement (captured from the
 * reference itself with `TACO_REF_DUMP=1 oracle/_ref/taco_ref_harness ...`), in particular:
 *   - operation order inside a row / fiber / slice (ascending position, scalar accumulator),
 *   - multiplication association ((B*C)*D for SDDMM and MTTKRP),
 *   - structure rules for sparse results (union keeps explicit zeros; SpGEMM keeps entries that sum to 0,
 *     columns ascending).
 * Compile with -ffp-contract=off so a*b+c is never fused: the reference's generated code is built by `cc`
 * for baseline x86-64 (no FMA), /root/reference/src/codegen/module.cpp:134-147.
 */
#include "taco_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}
int oracle_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void oracle_free(void* p) { free(p); }

static int cmp_i32(const void* a, const void* b) { /* reference prelude `cmp`, src/codegen/codegen_c.cpp:27-164 */
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

#define ORACLE_IMPL(T, S)                                                                                          \
  /* y(i) = A(i,j) * x(j), A CSR.  Default schedule: scalar accumulator, ascending position.                     \
     Loop nest: lowerForallPosition, src/lower/lowerer_impl_imperative.cpp:1304-1416 over                         \
     CompressedModeFormat::posIterBounds, src/lower/mode_format_compressed.cpp:80-86. */                          \
  void oracle_spmv_##S(int32_t n, const int32_t* pos, const int32_t* crd, const T* vals, const T* x, T* y) {      \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < n; i++) {                                \
      T acc = 0;                                                                                                   \
      for (int32_t p = pos[i]; p < pos[i + 1]; p++) acc += vals[p] * x[crd[p]];                                    \
      y[i] = acc;                                                                                                  \
    }                                                                                                              \
  }                                                                                                                \
  /* C(i,k) = A(i,j) * B(j,k), A CSR, B and C row-major dense.  zero C (initValues,                              \
     lowerer_impl_imperative.cpp:3348-3364) then C[i,k] = C[i,k] + A[p]*B[j,k], p ascending per row. */           \
  void oracle_spmm_##S(int32_t n, int32_t K, const int32_t* pos, const int32_t* crd, const T* vals, const T* B,   \
                       T* C) {                                                                                     \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < n; i++) {                                \
      T* c = C + (size_t)i * K;                                                                                    \
      for (int32_t k = 0; k < K; k++) c[k] = 0;                                                                    \
      for (int32_t p = pos[i]; p < pos[i + 1]; p++) {                                                              \
        const T* b = B + (size_t)crd[p] * K;                                                                       \
        T a = vals[p];                                                                                             \
        for (int32_t k = 0; k < K; k++) c[k] = c[k] + a * b[k];                                                    \
      }                                                                                                            \
    }                                                                                                              \
  }                                                                                                                \
  /* C(i,k) = A(i,j) * B(j,k), A doubly compressed {Sparse,Sparse} (the operand of the reference's spmmDCSRGPU test,  \
     test/tests-scheduling-eval.cpp:1309-1358).  Generated code (dumped through oracle/ref_harness.cpp): zero ALL of C,  \
     then for iA in [A1_pos[0], A1_pos[1]): i = A1_crd[iA]; for jA in [A2_pos[iA], A2_pos[iA+1]): j = A2_crd[jA];       \
     C[i,k] = C[i,k] + A[jA] * B[j,k] -- position loops of mode_format_compressed.cpp:80-105 at both levels. */         \
  void oracle_spmm_dcsr_##S(int32_t n, int32_t K, const int32_t* pos1, const int32_t* crd1, const int32_t* pos2,       \
                            const int32_t* crd2, const T* vals, const T* B, T* C) {                                    \
    _Pragma("omp parallel for schedule(static)") for (int64_t q = 0; q < (int64_t)n * K; q++) C[q] = 0;                \
    _Pragma("omp parallel for schedule(static)") for (int32_t iA = pos1[0]; iA < pos1[1]; iA++) {                      \
      T* c = C + (size_t)crd1[iA] * K;                                                                                 \
      for (int32_t p = pos2[iA]; p < pos2[iA + 1]; p++) {                                                              \
        const T* b = B + (size_t)crd2[p] * K;                                                                          \
        T a = vals[p];                                                                                                 \
        for (int32_t k = 0; k < K; k++) c[k] = c[k] + a * b[k];                                                        \
      }                                                                                                                \
    }                                                                                                                  \
  }                                                                                                                    \
  /* A(i,k) = B(i,k) * C(i,j) * D(j,k), A dense, D indexed (contraction, column): the statement of the reference's      \
     sddmmGPU test (test/tests-scheduling-eval.cpp:1360-1418).  Generated code (dumped through oracle/ref_harness.cpp):    \
     zero A; for i, for j (ascending), for kB in row i of B: A[i,k] = A[i,k] + (B[kB] * C[i,j]) * D[j,k]. */              \
  void oracle_sddmm_dense_##S(int32_t n, int32_t m, int32_t J, const int32_t* pos, const int32_t* crd, const T* Bvals,  \
                              const T* C, const T* D, T* A) {                                                          \
    _Pragma("omp parallel for schedule(static)") for (int64_t q = 0; q < (int64_t)n * m; q++) A[q] = 0;                \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < n; i++) {                                     \
      for (int32_t j = 0; j < J; j++) {                                                                                \
        const T cij = C[(size_t)i * J + j];                                                                            \
        for (int32_t p = pos[i]; p < pos[i + 1]; p++) {                                                                \
          const size_t a = (size_t)i * m + crd[p];                                                                     \
          A[a] = A[a] + (Bvals[p] * cij) * D[(size_t)j * m + crd[p]];                                                  \
        }                                                                                                              \
      }                                                                                                                \
    }                                                                                                                  \
  }                                                                                                                    \
  /* A(i,j) = B(i,j) * C(i,k) * D(j,k), A and B CSR (A has B's structure), C, D row-major.                       \
     tkA += (B[p] * C[i,k]) * D[j,k], k ascending, scalar accumulator; A_vals[jA++] = tkA. */                     \
  void oracle_sddmm_##S(int32_t n, int32_t K, const int32_t* pos, const int32_t* crd, const T* Bvals, const T* C, \
                        const T* D, T* Avals) {                                                                    \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < n; i++) {                                \
      const T* c = C + (size_t)i * K;                                                                              \
      for (int32_t p = pos[i]; p < pos[i + 1]; p++) {                                                              \
        const T* d = D + (size_t)crd[p] * K;                                                                       \
        T acc = 0;                                                                                                 \
        for (int32_t k = 0; k < K; k++) acc += (Bvals[p] * c[k]) * d[k];                                           \
        Avals[p] = acc;                                                                                            \
      }                                                                                                            \
    }                                                                                                              \
  }                                                                                                                \
  /* A(i,j) = B(i,k,l) * C(k,j) * D(l,j), B CSF {Compressed x3}, A, C, D row-major.                              \
     zero A; traversal i -> k -> l ascending; A[i,j] = A[i,j] + (B[p]*C[k,j])*D[l,j]. */                          \
  void oracle_mttkrp_##S(int32_t R, const int32_t* B1_pos, const int32_t* B1_crd, const int32_t* B2_pos,          \
                         const int32_t* B2_crd, const int32_t* B3_pos, const int32_t* B3_crd, const T* Bvals,     \
                         const T* C, const T* D, int32_t I, T* A) {                                                \
    memset(A, 0, sizeof(T) * (size_t)I * R);                                                                       \
    _Pragma("omp parallel for schedule(static)") for (int32_t iB = B1_pos[0]; iB < B1_pos[1]; iB++) {             \
      T* a = A + (size_t)B1_crd[iB] * R;                                                                           \
      for (int32_t kB = B2_pos[iB]; kB < B2_pos[iB + 1]; kB++) {                                                   \
        const T* c = C + (size_t)B2_crd[kB] * R;                                                                   \
        for (int32_t lB = B3_pos[kB]; lB < B3_pos[kB + 1]; lB++) {                                                 \
          const T* d = D + (size_t)B3_crd[lB] * R;                                                                 \
          T b = Bvals[lB];                                                                                         \
          for (int32_t j = 0; j < R; j++) a[j] = a[j] + (b * c[j]) * d[j];                                         \
        }                                                                                                          \
      }                                                                                                            \
    }                                                                                                              \
  }                                                                                                                \
  /* A(i,j) = B(i,j,k) * c(k): TTV, A dense I x K (reference GPU schedule scheduleTTVGPU,                        \
     test/tests-scheduling-eval.cpp:308-325).  Scalar accumulator per fiber, k ascending. */                      \
  void oracle_ttv_##S(const int32_t* B1_pos, const int32_t* B1_crd, const int32_t* B2_pos, const int32_t* B2_crd, \
                      const int32_t* B3_pos, const int32_t* B3_crd, const T* Bvals, const T* c, int32_t I,        \
                      int32_t K, T* A) {                                                                           \
    memset(A, 0, sizeof(T) * (size_t)I * K);                                                                       \
    _Pragma("omp parallel for schedule(static)") for (int32_t iB = B1_pos[0]; iB < B1_pos[1]; iB++) {             \
      int32_t i = B1_crd[iB];                                                                                      \
      for (int32_t jB = B2_pos[iB]; jB < B2_pos[iB + 1]; jB++) {                                                   \
        T acc = 0;                                                                                                 \
        for (int32_t kB = B3_pos[jB]; kB < B3_pos[jB + 1]; kB++) acc += Bvals[kB] * c[B3_crd[kB]];                 \
        A[(size_t)i * K + B2_crd[jB]] = acc;                                                                       \
      }                                                                                                            \
    }                                                                                                              \
  }                                                                                                                \
  /* A(i,j,l) = B(i,j,k) * C(k,l): TTM, A dense I x K x R (scheduleTTMGPU, :289-306). */                         \
  void oracle_ttm_##S(int32_t R, const int32_t* B1_pos, const int32_t* B1_crd, const int32_t* B2_pos,             \
                      const int32_t* B2_crd, const int32_t* B3_pos, const int32_t* B3_crd, const T* Bvals,        \
                      const T* C, int32_t I, int32_t K, T* A) {                                                    \
    memset(A, 0, sizeof(T) * (size_t)I * K * R);                                                                   \
    _Pragma("omp parallel for schedule(static)") for (int32_t iB = B1_pos[0]; iB < B1_pos[1]; iB++) {             \
      int32_t i = B1_crd[iB];                                                                                      \
      for (int32_t jB = B2_pos[iB]; jB < B2_pos[iB + 1]; jB++) {                                                   \
        T* a = A + ((size_t)i * K + B2_crd[jB]) * R;                                                               \
        for (int32_t kB = B3_pos[jB]; kB < B3_pos[jB + 1]; kB++) {                                                 \
          const T* cr = C + (size_t)B3_crd[kB] * R;                                                                \
          T b = Bvals[kB];                                                                                         \
          for (int32_t l = 0; l < R; l++) a[l] = a[l] + b * cr[l];                                                 \
        }                                                                                                          \
      }                                                                                                            \
    }                                                                                                              \
  }                                                                                                                \
  /* a(i,j) = A(i,k,j,l) * c(k,l): blocked SpMV, A = {Dense,Compressed,Dense,Dense} (the reference's `bspmv` test,  \
     test/tests-expr_storage.cpp:939-960).  Generated loop order i, kA, j, l; tl = sum_l A*c (scalar), a[i,j] = a[i,j] + tl  \
     after a zero fill -- captured with TACO_REF_DUMP=1 oracle/_ref/taco_ref_harness bspmv. */                                    \
  void oracle_bspmv_##S(int32_t Mb, int32_t br, int32_t bc, const int32_t* pos, const int32_t* crd, const T* vals,  \
                        const T* c, T* a) {                                                                          \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < Mb; i++) {                                 \
      T* ar = a + (size_t)i * br;                                                                                    \
      for (int32_t j = 0; j < br; j++) ar[j] = 0;                                                                    \
      for (int32_t kA = pos[i]; kA < pos[i + 1]; kA++) {                                                             \
        const T* cr = c + (size_t)crd[kA] * bc;                                                                      \
        for (int32_t j = 0; j < br; j++) {                                                                           \
          const T* blk = vals + ((size_t)kA * br + j) * bc;                                                          \
          T t = 0;                           /* scalar temporary per (block, row): tla_val in the generated C */    \
          for (int32_t l = 0; l < bc; l++) t += blk[l] * cr[l];                                                      \
          ar[j] = ar[j] + t;                                                                                         \
        }                                                                                                            \
      }                                                                                                              \
    }                                                                                                                \
  }                                                                                                                  \
  /* C(i,j,m) = A(i,k,j,l) * B(k,l,m): blocked SpMM, B = (Nb, bc, K) and C = (Mb, br, K) dense.  Generated loop      \
     order i, kA, j, l, m with C[mC] = C[mC] + A[lA] * B[mB] after a zero fill (dump of the reference, as above). */ \
  void oracle_bspmm_##S(int32_t Mb, int32_t br, int32_t bc, int32_t K, const int32_t* pos, const int32_t* crd,      \
                        const T* vals, const T* B, T* C) {                                                           \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < Mb; i++) {                                 \
      T* ci = C + (size_t)i * br * K;                                                                                \
      for (size_t q = 0; q < (size_t)br * K; q++) ci[q] = 0;                                                         \
      for (int32_t kA = pos[i]; kA < pos[i + 1]; kA++) {                                                             \
        const T* bk = B + (size_t)crd[kA] * bc * K;                                                                  \
        for (int32_t j = 0; j < br; j++) {                                                                           \
          T* cj = ci + (size_t)j * K;                                                                                \
          const T* blk = vals + ((size_t)kA * br + j) * bc;                                                          \
          for (int32_t l = 0; l < bc; l++) {                                                                         \
            const T av = blk[l];                                                                                     \
            const T* bl = bk + (size_t)l * K;                                                                        \
            for (int32_t m = 0; m < K; m++) cj[m] = cj[m] + av * bl[m];                                              \
          }                                                                                                          \
        }                                                                                                            \
      }                                                                                                              \
    }                                                                                                                \
  }                                                                                                                  \
  /* C(i,j) = A(i,j) + B(i,j): numeric phase; the same two-finger walk as assemble, values a+b | a | b           \
     (merge lattice union, src/lower/merge_lattice.cpp:1005; Appendix A.4 of SURVEY.md). */                       \
  void oracle_spadd_compute_##S(int32_t n, const int32_t* Apos, const int32_t* Acrd, const T* Avals,              \
                                const int32_t* Bpos, const int32_t* Bcrd, const T* Bvals, const int32_t* Cpos,    \
                                T* Cvals) {                                                                        \
    _Pragma("omp parallel for schedule(static)") for (int32_t i = 0; i < n; i++) {                                \
      int32_t a = Apos[i], ae = Apos[i + 1], b = Bpos[i], be = Bpos[i + 1], p = Cpos[i];                           \
      while (a < ae && b < be) {                                                                                   \
        int32_t ja = Acrd[a], jb = Bcrd[b], j = ja < jb ? ja : jb;                                                 \
        if (ja == j && jb == j) Cvals[p++] = Avals[a] + Bvals[b];                                                  \
        else if (ja == j) Cvals[p++] = Avals[a];                                                                   \
        else Cvals[p++] = Bvals[b];                                                                                \
        a += (ja == j);                                                                                            \
        b += (jb == j);                                                                                            \
      }                                                                                                            \
      while (a < ae) Cvals[p++] = Avals[a++];                                                                      \
      while (b < be) Cvals[p++] = Bvals[b++];                                                                      \
    }                                                                                                              \
  }                                                                                                                \
  /* C(i,k) = A(i,j) * B(j,k), all CSR: Gustavson with a dense row workspace (lowerWhere,                        \
     lowerer_impl_imperative.cpp:2516-2606): w[k] = a*b on first touch else w[k] = w[k] + a*b, in                 \
     (A-row order, then B-row order); drain in ascending column order; entries that sum to 0 are kept. */         \
  void oracle_spgemm_compute_##S(int32_t n, int32_t ncols, const int32_t* Apos, const int32_t* Acrd,              \
                                 const T* Avals, const int32_t* Bpos, const int32_t* Bcrd, const T* Bvals,        \
                                 const int32_t* Cpos, T* Cvals) {                                                  \
    _Pragma("omp parallel") {                                                                                      \
      T* w = (T*)malloc(sizeof(T) * (size_t)(ncols > 0 ? ncols : 1));                                              \
      uint8_t* set = (uint8_t*)calloc((size_t)(ncols > 0 ? ncols : 1), 1);                                         \
      int32_t* list = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ncols > 0 ? ncols : 1));                         \
      _Pragma("omp for schedule(dynamic, 64)") for (int32_t i = 0; i < n; i++) {                                  \
        int32_t sz = 0;                                                                                            \
        for (int32_t pa = Apos[i]; pa < Apos[i + 1]; pa++) {                                                       \
          int32_t j = Acrd[pa];                                                                                    \
          for (int32_t pb = Bpos[j]; pb < Bpos[j + 1]; pb++) {                                                     \
            int32_t k = Bcrd[pb];                                                                                  \
            if (!set[k]) { w[k] = Avals[pa] * Bvals[pb]; list[sz++] = k; set[k] = 1; }                             \
            else { w[k] = w[k] + Avals[pa] * Bvals[pb]; }                                                          \
          }                                                                                                        \
        }                                                                                                          \
        qsort(list, (size_t)sz, sizeof(int32_t), cmp_i32);                                                         \
        int32_t p = Cpos[i];                                                                                       \
        for (int32_t q = 0; q < sz; q++) { Cvals[p++] = w[list[q]]; set[list[q]] = 0; }                            \
      }                                                                                                            \
      free(w); free(set); free(list);                                                                              \
    }                                                                                                              \
  }

ORACLE_IMPL(float, f32)
ORACLE_IMPL(double, f64)

/* SDDMM with CSR result: append assembly (initResultArrays / getAppendCoord / finalizeResultArrays,
 * lowerer_impl_imperative.cpp:3157-3308; prefix sum mode_format_compressed.cpp:192-211) -- the result has exactly
 * B's structure. */
void oracle_sddmm_assemble(int32_t n, const int32_t* Bpos, const int32_t* Bcrd, int32_t* Apos, int32_t** Acrd) {
  int32_t nnz = Bpos[n] - Bpos[0];
  int32_t* crd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  int32_t jA = 0;
  Apos[0] = 0;
  for (int32_t i = 0; i < n; i++) {
    int32_t begin = jA;
    for (int32_t p = Bpos[i]; p < Bpos[i + 1]; p++) crd[jA++] = Bcrd[p];
    Apos[i + 1] = jA - begin;
  }
  for (int32_t i = 0; i < n; i++) Apos[i + 1] += Apos[i];
  *Acrd = crd;
}

/* SpAdd two-phase Insert assembly (lowerAssemble :2616-2779 + CompressedModeFormat getSeqInitEdges/
 * getSeqInsertEdge/getYieldPos/getFinalizeYieldPos, mode_format_compressed.cpp:217-271): symbolic union count,
 * exclusive scan, crd fill ascending.  Identical final structure to the default append strategy. */
void oracle_spadd_assemble(int32_t n, const int32_t* Apos, const int32_t* Acrd, const int32_t* Bpos,
                           const int32_t* Bcrd, int32_t* Cpos, int32_t** Ccrd) {
  int32_t* nnz = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < n; i++) {
    int32_t a = Apos[i], ae = Apos[i + 1], b = Bpos[i], be = Bpos[i + 1], c = 0;
    while (a < ae && b < be) {
      int32_t ja = Acrd[a], jb = Bcrd[b], j = ja < jb ? ja : jb;
      c++;
      a += (ja == j);
      b += (jb == j);
    }
    nnz[i] = c + (ae - a) + (be - b);
  }
  Cpos[0] = 0;
  for (int32_t i = 0; i < n; i++) Cpos[i + 1] = Cpos[i] + nnz[i];
  free(nnz);
  int32_t* crd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(Cpos[n] > 0 ? Cpos[n] : 1));
#pragma omp parallel for schedule(static)
  for (int32_t i = 0; i < n; i++) {
    int32_t a = Apos[i], ae = Apos[i + 1], b = Bpos[i], be = Bpos[i + 1], p = Cpos[i];
    while (a < ae && b < be) {
      int32_t ja = Acrd[a], jb = Bcrd[b], j = ja < jb ? ja : jb;
      crd[p++] = j;
      a += (ja == j);
      b += (jb == j);
    }
    while (a < ae) crd[p++] = Acrd[a++];
    while (b < be) crd[p++] = Bcrd[b++];
  }
  *Ccrd = crd;
}

/* SpGEMM two-phase assembly (Appendix A.5): symbolic count of reachable columns per row, scan, sorted crd. */
void oracle_spgemm_assemble(int32_t n, int32_t ncols, const int32_t* Apos, const int32_t* Acrd, const int32_t* Bpos,
                            const int32_t* Bcrd, int32_t* Cpos, int32_t** Ccrd) {
  int32_t* nnz = (int32_t*)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t));
  size_t nc = (size_t)(ncols > 0 ? ncols : 1);
#pragma omp parallel
  {
    uint8_t* set = (uint8_t*)calloc(nc, 1);
    int32_t* list = (int32_t*)malloc(sizeof(int32_t) * nc);
#pragma omp for schedule(dynamic, 64)
    for (int32_t i = 0; i < n; i++) {
      int32_t sz = 0;
      for (int32_t pa = Apos[i]; pa < Apos[i + 1]; pa++) {
        int32_t j = Acrd[pa];
        for (int32_t pb = Bpos[j]; pb < Bpos[j + 1]; pb++) {
          int32_t k = Bcrd[pb];
          if (!set[k]) { list[sz++] = k; set[k] = 1; }
        }
      }
      for (int32_t q = 0; q < sz; q++) set[list[q]] = 0;
      nnz[i] = sz;
    }
    free(set); free(list);
  }
  Cpos[0] = 0;
  for (int32_t i = 0; i < n; i++) Cpos[i + 1] = Cpos[i] + nnz[i];
  free(nnz);
  int32_t* crd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(Cpos[n] > 0 ? Cpos[n] : 1));
#pragma omp parallel
  {
    uint8_t* set = (uint8_t*)calloc(nc, 1);
    int32_t* list = (int32_t*)malloc(sizeof(int32_t) * nc);
#pragma omp for schedule(dynamic, 64)
    for (int32_t i = 0; i < n; i++) {
      int32_t sz = 0;
      for (int32_t pa = Apos[i]; pa < Apos[i + 1]; pa++) {
        int32_t j = Acrd[pa];
        for (int32_t pb = Bpos[j]; pb < Bpos[j + 1]; pb++) {
          int32_t k = Bcrd[pb];
          if (!set[k]) { list[sz++] = k; set[k] = 1; }
        }
      }
      qsort(list, (size_t)sz, sizeof(int32_t), cmp_i32);
      int32_t p = Cpos[i];
      for (int32_t q = 0; q < sz; q++) { crd[p++] = list[q]; set[list[q]] = 0; }
    }
    free(set); free(list);
  }
  *Ccrd = crd;
}
