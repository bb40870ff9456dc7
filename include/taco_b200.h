/* taco_b200.h -- C ABI of libtaco_b200.so, the B200-native (sm_100a) GPU compute path for taco.
 *
 * This is the drop-in boundary.  The reference JIT-compiles one shared object per
 * concrete index statement and calls it through `taco_tensor_t*`:
 *     int compute (taco_tensor_t* result..., taco_tensor_t* operand...)     src/codegen/codegen.cpp:514-585
 *     int assemble(...same...)   int evaluate(...same...)
 *     extern "C" int _shim_<name>(void** parameterPack)                      src/codegen/codegen_cuda.cpp:1500-1540
 *   reached via Module::callFuncPacked(name, void** args)                    include/taco/codegen/module.h:57-64
 * libtaco_b200.so exports, for every kernel family on the hot path, exactly those entry points (prefixed with
 * the family name because one library holds all families), plus a module object that resolves
 * (index expression, formats, dtype) to a family the way Module::compile()+getFuncPtr() would
 * (src/codegen/module.cpp:111-180).  Argument order is the reference's: results first, then operands in order of
 * first appearance in the statement (src/tensor.cpp:778-806).  Return value 0 = success (codegen_cuda.cpp:806);
 * non-zero = error, message via taco_b200_last_error() -- the library never calls exit() (the reference's
 * gpuErrchk does, codegen_cuda.cpp:58-67).
 *
 * Memory contract
 *   - Every array pointer inside a taco_tensor_t (indices[l][k], vals) may be a HOST pointer (pageable or
 *     pinned) or a DEVICE pointer; each is classified per call.  Host operands are staged to HBM (pinned
 *     double-buffered for pageable memory, direct DMA for pinned memory) unless the array is registered resident
 *     (taco_b200_make_resident); device operands are used in place with zero copies.
 *   - assemble() allocates result arrays (dense vals; pos/crd/vals of sparse results) in the result space chosen
 *     with taco_b200_set_result_space(): HOST (default; malloc()ed, so taco's Array::Free policy can free() them,
 *     src/tensor.cpp:278-290, src/storage/array.cpp:22-37) or DEVICE (cudaMalloc; release with taco_b200_free).
 *   - compute() writes result values where result->vals points (host or device).
 *   - Work is enqueued on the stream set by taco_b200_set_stream() (default: a private non-blocking stream).
 *     Entry points with host-visible results synchronise that stream before returning; with all-device tensors
 *     they return immediately (stream-ordered), like a kernel launch.  Device operands written by, and device results
 *     read by, work on ANOTHER stream are the caller's to order: hand the library that stream (taco_b200_set_stream)
 *     or synchronise (taco_b200_synchronize) -- exactly as between any two CUDA streams.
 *
 * There is NO CPU fallback: every entry point fails with TACO_B200_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef TACO_B200_H
#define TACO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Layout-identical to the reference's struct (include/taco/taco_tensor_t.h:11-23); the include guard is the
 * reference's own so both headers can be included together. */
#ifndef TACO_TENSOR_T_DEFINED
#define TACO_TENSOR_T_DEFINED
typedef enum { taco_mode_dense, taco_mode_sparse } taco_mode_t;
/* Convention for DEVICE-RESIDENT sparse tensors only: vals_size holds the number of stored values (nnz), so the
 * library never has to read pos[] back to size a launch.  For host-described tensors vals_size is ignored (the
 * reference leaves it uninitialised, src/taco_tensor_t.cpp:32-67) and nnz is read from pos[] on the host. */
typedef struct taco_tensor_t {
  int32_t      order;         /* number of modes */
  int32_t*     dimensions;    /* [order], indexed by MODE */
  int32_t      csize;         /* component size in BITS (src/storage/storage.cpp:61) */
  int32_t*     mode_ordering; /* [order]: storage level l stores mode mode_ordering[l] */
  taco_mode_t* mode_types;    /* [order], per storage level */
  uint8_t***   indices;       /* [order][k]: dense level -> [0]=int32[1]{dim}; compressed -> [0]=pos, [1]=crd */
  uint8_t*     vals;          /* values, level order */
  uint8_t*     fill_value;
  int32_t      vals_size;
} taco_tensor_t;
#endif

/* ---- status ------------------------------------------------------------------------------------------- */
#define TACO_B200_OK              0
#define TACO_B200_ERR_CUDA        1   /* CUDA runtime error or no usable device */
#define TACO_B200_ERR_FORMAT      2   /* tensor format / dtype not handled by this kernel family */
#define TACO_B200_ERR_ARG         3   /* bad argument (NULL, dimension mismatch, int32 overflow ...) */
#define TACO_B200_ERR_UNSUPPORTED 4   /* statement is not on the GPU hot path (no CPU fallback exists) */
#define TACO_B200_ERR_ALLOC       5

const char* taco_b200_last_error(void);          /* thread-local message of the last failing call */
const char* taco_b200_version(void);

/* ---- runtime (replaces src/cuda.cpp: cuda_unified_alloc/free, should_use_CUDA_*) ---------------------- */
int  taco_b200_init(int device);                  /* idempotent; selects the device for this process */
int  taco_b200_device_count(void);
int  taco_b200_set_stream(void* cuda_stream);     /* cudaStream_t; NULL restores the private stream; pass
                                                     cudaStreamLegacy (0x1) for the legacy default stream */
void* taco_b200_get_stream(void);
int  taco_b200_synchronize(void);
int  taco_b200_launch_count(void);                /* number of kernels launched by this library so far */

/* Per-kernel device timing for bench.py's roofline line: when enabled, each dominant kernel launch is bracketed by
 * CUDA events on the launch stream.  kernel_name in {"spmv_csr","spmm_csr","sddmm_csr","mttkrp_csf","ttv_csf",
 * "ttm_csf","bspmv_bcsr","bspmm_bcsr","spadd_symbolic","spadd_numeric","spgemm_symbolic","spgemm_numeric"}. */
int  taco_b200_profile_enable(int on);
int  taco_b200_profile_get(const char* kernel_name, double* total_ms, int* launches);
int  taco_b200_profile_reset(void);

#define TACO_B200_SPACE_HOST   0
#define TACO_B200_SPACE_DEVICE 1
int  taco_b200_set_result_space(int space);
int  taco_b200_get_result_space(void);

void* taco_b200_host_alloc(size_t bytes);         /* pinned host memory (fast staging) */
void  taco_b200_host_free(void* p);
void* taco_b200_device_alloc(size_t bytes);
void  taco_b200_free(void* p);                    /* frees device OR pinned-host memory from this library */

/* Fused compute + all-gather over NVLink (multi-GPU).  Register a window of THIS GPU's memory together with
 * the NVLink multicast mapping of the same symmetric allocation (e.g. torch.distributed._symmetric_memory: buffer_ptrs[rank] and
 * multicast_ptr).  A dense result (SpMM C, MTTKRP A) that lies inside the window is then stored through the multicast address:
 * every row a kernel produces is delivered by the NVSwitch to all GPUs of the group in the same store, so after the step (and
 * a cross-rank barrier) every rank holds the whole gathered result -- no separate all-gather.  NULL, NULL, 0 clears. */
int  taco_b200_set_result_multicast(const void* local_base, void* multicast_base, size_t bytes);
/* The same fusion with plain peer-to-peer stores: `peer_bases[i]` is the window of peer GPU i as mapped into THIS process
 * (symmetric memory buffer_ptrs of the other ranks, or cudaIpcOpenMemHandle / cudaDeviceEnablePeerAccess pointers), 1 <= npeers <= 7.
 * Every result row is stored locally and to each peer from inside the kernel.  Each GPU then receives N-1 shards over NVLink
 * instead of the N a multicast store delivers (the switch also loops the sender's own copy back): the better trade for small
 * groups.  Replaces a registered multicast window and vice versa; NULL / npeers 0 clears.  Register or clear a window only
 * while no compute call of this process is in flight. */
int  taco_b200_set_result_peers(const void* local_base, size_t bytes, int npeers, void* const* peer_bases);

/* Residency cache: keep a device mirror of an immutable host array across calls (pinned upload once). */
int  taco_b200_make_resident(const void* host_ptr, size_t bytes);
int  taco_b200_invalidate(const void* host_ptr);  /* host array changed or is about to be freed */
int  taco_b200_drop_all_resident(void);

/* ---- kernel families: <family>_{assemble,compute,evaluate} + _shim_ variants -------------------------- */
/* y(i) = A(i,j) * x(j)            A CSR {Dense,Compressed}; x, y dense; fp32 | fp64
 * replaces the kernels emitted for scheduleSpMVGPU / SplitPosGPU / RowsGPU (test/tests-scheduling-eval.cpp:193-247) */
int taco_b200_spmv_assemble(taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x);
int taco_b200_spmv_compute (taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x);
int taco_b200_spmv_evaluate(taco_tensor_t* y, taco_tensor_t* A, taco_tensor_t* x);

/* C(i,k) = A(i,j) * B(j,k)        A CSR; B dense row-major; C dense row-major or {Dense,Dense},{1,0} (the
 * reference GPU test's column-major C, tests-scheduling-eval.cpp:1258-1307); fp32 | fp64
 * replaces scheduleSpMMGPU (:249-268) */
int taco_b200_spmm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spmm_compute (taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spmm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);

/* A(i,j) = B(i,j) * C(i,k) * D(j,k)   A, B CSR (A gets B's structure); C, D dense row-major
 * replaces scheduleSDDMMGPU (:270-287) (CSR-output form) */
int taco_b200_sddmm_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);
int taco_b200_sddmm_compute (taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);
int taco_b200_sddmm_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);

/* A(i,k) = B(i,k) * C(i,j) * D(j,k)   B CSR; A, C, D dense row-major, D indexed (contraction, column): the statement of the
 * reference's sddmmGPU test (test/tests-scheduling-eval.cpp:1360-1418, dense result); replaces scheduleSDDMMGPU (:270-287) */
int taco_b200_sddmm_dense_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);
int taco_b200_sddmm_dense_compute (taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);
int taco_b200_sddmm_dense_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);

/* A(i,j) = B(i,k,l) * C(k,j) * D(l,j)   B CSF {Compressed x3}; A, C, D dense row-major
 * replaces scheduleMTTKRPGPU (:327-342) */
int taco_b200_mttkrp_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);
int taco_b200_mttkrp_compute (taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);
int taco_b200_mttkrp_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C, taco_tensor_t* D);

/* C(i,k) = A(i,j) * B(j,k)   A = {Compressed,Compressed} (doubly compressed rows), B, C dense: the statement of the
 * reference's spmmDCSRGPU test (test/tests-scheduling-eval.cpp:1309-1358); replaces what CodeGen_CUDA emits for
 * scheduleSpMMNZRowsGPU (:358-369).  The level-0 row list is expanded to a CSR pos array on the device and the CSR
 * kernel runs on it; rows that are not stored come out as zero rows (the reference zero-fills C). */
int taco_b200_spmm_dcsr_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spmm_dcsr_compute (taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spmm_dcsr_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);

/* a(i,j) = A(i,k,j,l) * c(k,l)  (blocked SpMV)   and   C(i,j,m) = A(i,k,j,l) * B(k,l,m)  (blocked SpMM)
 * A = {Dense,Compressed,Dense,Dense} over (block row, block column, row in block, column in block), the blocked
 * format of the reference's `bspmv` test (test/tests-expr_storage.cpp:939-960); c, a, B, C dense row-major; fp32 | fp64.
 * Blocked SpMM is the one hot-path contraction that is dense inside a block: fp32 with 16x16 / 32x32 blocks runs on the
 * tcgen05 tensor cores (3xTF32 split, fp32 accumulation in tensor memory); everything else keeps the reference's
 * summation order on CUDA cores.  Replaces what CodeGen_CUDA emits for these statements (codegen_cuda.cpp:619-812). */
int taco_b200_bspmv_assemble(taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c);
int taco_b200_bspmv_compute (taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c);
int taco_b200_bspmv_evaluate(taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c);
int taco_b200_bspmm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_bspmm_compute (taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_bspmm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);

/* A(i,j) = B(i,j,k) * c(k)  (TTV)   and   A(i,j,l) = B(i,j,k) * C(k,l)  (TTM);  B CSF, A dense
 * replace scheduleTTVGPU (:308-325) and scheduleTTMGPU (:289-306) */
int taco_b200_ttv_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* c);
int taco_b200_ttv_compute (taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* c);
int taco_b200_ttv_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* c);
int taco_b200_ttm_assemble(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C);
int taco_b200_ttm_compute (taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C);
int taco_b200_ttm_evaluate(taco_tensor_t* A, taco_tensor_t* B, taco_tensor_t* C);

/* C(i,j) = A(i,j) + B(i,j)        all CSR.  GPU two-phase assembly (symbolic count -> exclusive scan -> crd fill),
 * the device version of lowerAssemble + CompressedModeFormat getSeqInsertEdge/getYieldPos
 * (src/lower/lowerer_impl_imperative.cpp:2616-2779, src/lower/mode_format_compressed.cpp:217-271).
 * pos/crd bit-exact with the reference, explicit zeros kept. */
int taco_b200_spadd_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spadd_compute (taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spadd_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);

/* C(i,k) = A(i,j) * B(j,k)        all CSR (Gustavson; lowerWhere workspace semantics, :2516-2606): columns
 * ascending, entries that sum to zero kept. */
int taco_b200_spgemm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spgemm_compute (taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);
int taco_b200_spgemm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B);

/* COO -> level arrays on the device: the replacement for TensorBase::pack()'s host qsort + JIT-compiled `pack` helper
 * (src/tensor.cpp:295-463; helper signature `int pack(taco_tensor_t* A, taco_tensor_t* B)`, src/tensor.cpp:932-1000).
 * `coo` is the reference's coordinate-buffer tensor: every level "sparse", indices[0][0] = int32{0, n}, indices[l][1] =
 * the n coordinates of level l, vals = the n components (host or device arrays; ANY order -- the sort happens here).
 * `A` gives the target format -- {Dense,Compressed}, {Compressed,Compressed} or {Compressed,Compressed,Compressed}, any
 * mode ordering (levels in storage order on both sides, e.g. {1,0} = CSC) -- and receives freshly allocated pos / crd / vals in the configured result space, exactly as
 * an assemble call does; A->vals_size = number of stored components.  Equal coordinates are added (insertion order). */
int taco_b200_pack(taco_tensor_t* A, taco_tensor_t* coo);
int _shim_taco_b200_pack(void** p);

/* File -> packed tensor, parsed on the device: the replacement for taco::read(filename, format) on the two text formats the
 * reference's tests and CLI use -- Matrix Market coordinate files (.mtx / .ttx, `real`, `general` | `symmetric`;
 * src/storage/file_io_mtx.cpp:39-150) and FROSTT files (.tns; src/storage/file_io_tns.cpp:39-96).  The bytes are uploaded once,
 * split into lines and parsed by one thread per entry (values bit-identical to strtod), then packed by taco_b200_pack.
 * `A` gives the target format and component type as for taco_b200_pack; A->dimensions[m] <= 0 means "take it from the file"
 * (the size line of a .mtx, the largest coordinate of a .tns -- as the reference infers it) and is filled in. */
int taco_b200_read(const char* path, taco_tensor_t* A);
int _shim_taco_b200_read(void** p);       /* p[0] = const char* path, p[1] = taco_tensor_t* */

/* ---- module object: the replacement for ir::Module on the GPU path ------------------------------------ */
/* expr    : index notation as the CLI / Tensor API prints it, e.g. "y(i) = A(i,j) * x(j)"
 * formats : comma separated per-tensor level formats in CLI syntax (tools/taco.cpp -f=), e.g. "A:ds,x:d,y:d";
 *           an optional mode ordering follows a second colon ("C:dd:1,0").  Tensors not listed are dense.
 * dtype   : "f32" | "f64"
 * Returns NULL (and sets last_error, TACO_B200_ERR_UNSUPPORTED) when the statement is not a recognised hot-path
 * pattern -- there is no CPU fallback. */
typedef struct taco_b200_module taco_b200_module_t;
taco_b200_module_t* taco_b200_module_open(const char* expr, const char* formats, const char* dtype);
/* Same, with the order in which the caller will pack the tensors given explicitly (comma separated names; NULL = results
 * first, then operands by first appearance in `expr` -- taco's own order, src/tensor.cpp:778-806).  The operands of a product
 * or a sum commute: `y(i) = x(j) * A(i,j)` opens the spmv family and call_packed / the stub source reorder the pack. */
taco_b200_module_t* taco_b200_module_open_args(const char* expr, const char* formats, const char* dtype, const char* args);
const char* taco_b200_module_family(const taco_b200_module_t* m);   /* "spmv", "spmm", ... */
int   taco_b200_module_num_args(const taco_b200_module_t* m);
/* name in {"assemble","compute","evaluate"}; args = packed taco_tensor_t* exactly as Module::callFuncPacked */
int   taco_b200_module_call_packed(taco_b200_module_t* m, const char* name, void** args);
void* taco_b200_module_get_func_ptr(taco_b200_module_t* m, const char* name);  /* like Module::getFuncPtr */
void  taco_b200_module_close(taco_b200_module_t* m);
/* C source for the reference's own plug point TensorBase::compileSource(std::string) (src/tensor.cpp:905-930) and the
 * CLI's -read-source= (tools/taco.cpp:1212-1259): defines assemble/compute/evaluate with the signatures taco's
 * generated shims call (src/codegen/codegen_c.cpp:591-628) and forwards them to this library (dlopen of
 * $TACO_B200_LIB, else "libtaco_b200.so").  With it an UNMODIFIED taco runs the statement on the GPU path. */
const char* taco_b200_module_stub_source(taco_b200_module_t* m);

/* _shim_ entry points with the reference's exact shim signature (codegen_cuda.cpp:1500-1540). */
int _shim_taco_b200_spmv_assemble(void** p);   int _shim_taco_b200_spmv_compute(void** p);   int _shim_taco_b200_spmv_evaluate(void** p);
int _shim_taco_b200_spmm_assemble(void** p);   int _shim_taco_b200_spmm_compute(void** p);   int _shim_taco_b200_spmm_evaluate(void** p);
int _shim_taco_b200_spmm_dcsr_assemble(void** p); int _shim_taco_b200_spmm_dcsr_compute(void** p); int _shim_taco_b200_spmm_dcsr_evaluate(void** p);
int _shim_taco_b200_sddmm_assemble(void** p);  int _shim_taco_b200_sddmm_compute(void** p);  int _shim_taco_b200_sddmm_evaluate(void** p);
int _shim_taco_b200_sddmm_dense_assemble(void** p); int _shim_taco_b200_sddmm_dense_compute(void** p); int _shim_taco_b200_sddmm_dense_evaluate(void** p);
int _shim_taco_b200_mttkrp_assemble(void** p); int _shim_taco_b200_mttkrp_compute(void** p); int _shim_taco_b200_mttkrp_evaluate(void** p);
int _shim_taco_b200_ttv_assemble(void** p);    int _shim_taco_b200_ttv_compute(void** p);    int _shim_taco_b200_ttv_evaluate(void** p);
int _shim_taco_b200_ttm_assemble(void** p);    int _shim_taco_b200_ttm_compute(void** p);    int _shim_taco_b200_ttm_evaluate(void** p);
int _shim_taco_b200_bspmv_assemble(void** p);  int _shim_taco_b200_bspmv_compute(void** p);  int _shim_taco_b200_bspmv_evaluate(void** p);
int _shim_taco_b200_bspmm_assemble(void** p);  int _shim_taco_b200_bspmm_compute(void** p);  int _shim_taco_b200_bspmm_evaluate(void** p);
int _shim_taco_b200_spadd_assemble(void** p);  int _shim_taco_b200_spadd_compute(void** p);  int _shim_taco_b200_spadd_evaluate(void** p);
int _shim_taco_b200_spgemm_assemble(void** p); int _shim_taco_b200_spgemm_compute(void** p); int _shim_taco_b200_spgemm_evaluate(void** p);

/* ---- multi-GPU partitioner (new; the reference has none -- SURVEY.md section 8(e)) -------------------- */
/* Split `parent_size` pos-ranges into `parts` contiguous chunks balanced by child count:
 * bounds[g] = smallest r with pos[r] >= g*pos[parent_size]/parts, bounds[0]=0, bounds[parts]=parent_size.
 * `pos` may be a host or device pointer; `bounds` is host int32[parts+1].  (Same search as the reference's
 * taco_binarySearchBeforeBlock, codegen_cuda.cpp:110-125, applied across GPUs.) */
int taco_b200_partition_pos(const int32_t* pos, int32_t parent_size, int32_t parts, int32_t* bounds);

#ifdef __cplusplus
}
#endif
#endif /* TACO_B200_H */
