/* oracle/taco_oracle.h -- TEST INFRASTRUCTURE ONLY (the CHECKER, never the thing shipped or measured as product).
 *
 * Plain-C restatement of the kernels the reference (tensor-compiler/taco) GENERATES for the hot-path
 * statements, i.e. the output of /root/reference/src/codegen/codegen_c.cpp for the loop nests built by
 * /root/reference/src/lower/lowerer_impl_imperative.cpp.  Each function cites what it follows.
 *
 * Parity pin: tests/test_oracle_golden.py checks every function against
 *   (1) the reference's own known-answer vectors (test/tests-expr_storage.cpp groups spmv :873, matrix_add :523,
 *       matrix_mul :996, bspmv :939, tensor_vector_mul :1037, tensor_matrix_mul :1096, mttkrp :1145 over the fixtures in
 *       test/test_tensors.cpp), and
 *   (2) outputs of the reference itself (oracle/_ref/taco_ref_harness, built from /root/reference by
 *       oracle/Makefile) committed as fixtures under tests/golden/ by tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#ifndef TACO_ORACLE_H
#define TACO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

void oracle_set_num_threads(int n);
int  oracle_get_max_threads(void);

#define ORACLE_DECL(T, S)                                                                                          \
  void oracle_spmv_##S(int32_t n, const int32_t* pos, const int32_t* crd, const T* vals, const T* x, T* y);       \
  void oracle_spmm_##S(int32_t n, int32_t K, const int32_t* pos, const int32_t* crd, const T* vals, const T* B,   \
                       T* C);                                                                                      \
  void oracle_spmm_dcsr_##S(int32_t n, int32_t K, const int32_t* pos1, const int32_t* crd1, const int32_t* pos2,      \
                            const int32_t* crd2, const T* vals, const T* B, T* C);                                    \
  void oracle_sddmm_dense_##S(int32_t n, int32_t m, int32_t J, const int32_t* pos, const int32_t* crd, const T* Bvals,  \
                              const T* C, const T* D, T* A);                                                           \
  void oracle_sddmm_##S(int32_t n, int32_t K, const int32_t* pos, const int32_t* crd, const T* Bvals, const T* C, \
                        const T* D, T* Avals);                                                                     \
  void oracle_mttkrp_##S(int32_t R, const int32_t* B1_pos, const int32_t* B1_crd, const int32_t* B2_pos,          \
                         const int32_t* B2_crd, const int32_t* B3_pos, const int32_t* B3_crd, const T* Bvals,     \
                         const T* C, const T* D, int32_t I, T* A);                                                 \
  void oracle_ttv_##S(const int32_t* B1_pos, const int32_t* B1_crd, const int32_t* B2_pos, const int32_t* B2_crd, \
                      const int32_t* B3_pos, const int32_t* B3_crd, const T* Bvals, const T* c, int32_t I,        \
                      int32_t K, T* A);                                                                            \
  void oracle_ttm_##S(int32_t R, const int32_t* B1_pos, const int32_t* B1_crd, const int32_t* B2_pos,             \
                      const int32_t* B2_crd, const int32_t* B3_pos, const int32_t* B3_crd, const T* Bvals,        \
                      const T* C, int32_t I, int32_t K, T* A);                                                     \
  void oracle_bspmv_##S(int32_t Mb, int32_t br, int32_t bc, const int32_t* pos, const int32_t* crd, const T* vals,  \
                        const T* c, T* a);                                                                           \
  void oracle_bspmm_##S(int32_t Mb, int32_t br, int32_t bc, int32_t K, const int32_t* pos, const int32_t* crd,      \
                        const T* vals, const T* B, T* C);                                                            \
  void oracle_spadd_compute_##S(int32_t n, const int32_t* Apos, const int32_t* Acrd, const T* Avals,              \
                                const int32_t* Bpos, const int32_t* Bcrd, const T* Bvals, const int32_t* Cpos,    \
                                T* Cvals);                                                                         \
  void oracle_spgemm_compute_##S(int32_t n, int32_t ncols, const int32_t* Apos, const int32_t* Acrd,              \
                                 const T* Avals, const int32_t* Bpos, const int32_t* Bcrd, const T* Bvals,        \
                                 const int32_t* Cpos, T* Cvals);

ORACLE_DECL(float, f32)
ORACLE_DECL(double, f64)

/* structure-only phases (dtype independent).  Cpos has n+1 entries; *Ccrd is malloc()ed, caller frees. */
void oracle_sddmm_assemble(int32_t n, const int32_t* Bpos, const int32_t* Bcrd, int32_t* Apos, int32_t** Acrd);
void oracle_spadd_assemble(int32_t n, const int32_t* Apos, const int32_t* Acrd, const int32_t* Bpos,
                           const int32_t* Bcrd, int32_t* Cpos, int32_t** Ccrd);
void oracle_spgemm_assemble(int32_t n, int32_t ncols, const int32_t* Apos, const int32_t* Acrd, const int32_t* Bpos,
                            const int32_t* Bcrd, int32_t* Cpos, int32_t** Ccrd);
void oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
