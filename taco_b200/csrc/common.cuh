// common.cuh -- shared declarations of libtaco_b200 (runtime context, tensor views, device helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <mutex>
#include <string>

#include "../../include/taco_b200.h"

namespace tb {

// ---------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);   // records thread-local message, returns code
#define TB_CUDA(expr)                                                                             \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return tb::fail(TACO_B200_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
#define TB_TRY(expr)                    \
  do {                                  \
    int _rc = (expr);                   \
    if (_rc != TACO_B200_OK) return _rc; \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
// runtime context (runtime.cu)
// ---------------------------------------------------------------------------------------------------------
int ensure_init();
cudaStream_t stream();
cudaStream_t aux_stream(int i);                        // 0: upload, 1: download (host-operand pipelines), 2: side compute stream
bool is_resident(const void* host_ptr, size_t bytes);  // registered with taco_b200_make_resident
int num_sms();
size_t device_mem_total();
void count_launch(int n = 1);
int result_space();

enum class Mem { Host, Pinned, Device };
Mem classify(const void* p);
// true only for plain device memory (cudaMalloc / pool): the one case where the vals_size convention of taco_b200.h applies.
// Managed memory classifies as Device (the kernels can dereference it) but is what the REFERENCE allocates with unified
// memory on, and the reference never initialises vals_size (src/taco_tensor_t.cpp:32-67): sizes are read from pos[] there.
bool trusts_vals_size(const void* p);

// A read-only operand array made visible to the device for the duration of one call.
// dev() is usable on stream() after acquire(); release happens in the destructor (stream-ordered free).
struct In {
  const void* dptr = nullptr;
  void* owned = nullptr;    // stream-ordered scratch to free
  ~In();
  int acquire(const void* p, size_t bytes);
  template <typename T> const T* as() const { return (const T*)dptr; }
};

// A result array: device buffer the kernels write, copied back to the caller's pointer if that is host memory.
struct Out {
  void* dptr = nullptr;
  void* owned = nullptr;
  void* host_dst = nullptr;
  size_t bytes = 0;
  ~Out();
  int acquire(void* p, size_t bytes);      // p: where the caller wants the result (host or device)
  int commit();                            // enqueue D2H if needed; sets need_sync()
  template <typename T> T* as() const { return (T*)dptr; }
};
bool need_sync();          // a host-visible result was produced in this call
void clear_need_sync();
int finish_call();         // synchronise the stream iff need_sync()

// optional per-kernel timing (taco_b200_profile_enable): CUDA events recorded on the launch stream around a kernel
struct ProfScope {
  int slot = -1;
  explicit ProfScope(const char* kernel_name);
  ~ProfScope();
};

// Where a dense result row goes besides (or instead of) its local address -- the all-gather of the result fused into the
// kernel that produces it (taco_b200_set_result_multicast / taco_b200_set_result_peers).  result_fanout() returns a non-empty
// fan-out when the result [p, p + bytes) lies inside the registered local window:
//   n == -1: ONE store through the NVLink multicast mapping, d[0] bytes away: the NVSwitch delivers it to every GPU of the
//            group, the local one included (each GPU receives N shards over NVLink);
//   n  >  0: the local store plus n peer-to-peer stores, d[i] bytes away (the same window of peer i, mapped into this
//            process): each GPU receives N-1 shards and sends n copies of its own -- the better trade for small groups.
struct Fanout {
  int n = 0;
  long long d[7] = {0, 0, 0, 0, 0, 0, 0};
};
Fanout result_fanout(const void* p, size_t bytes);

// A persistent 256-byte device buffer for tiny synchronous reductions (counters read back right away).  Deliberately NOT a
// pool allocation: a 16-byte cudaMallocAsync between the large result arrays of consecutive calls splits the pool's big free
// blocks, and every other call then pays milliseconds for fresh physical memory (measured on SpAdd: 0.25 -> 0.6-6 ms).
// The caller holds small_scratch_mutex() from the first use until its read-back has completed.
void* small_scratch();
std::mutex& small_scratch_mutex();

// stream-ordered scratch
int scratch_alloc(void** p, size_t bytes);
void scratch_free(void* p);
// owns the scratch buffers and events of a multi-stream pipeline: released on every exit path (early error returns included)
struct PipelineGuard {
  void* bufs[32];
  cudaEvent_t evs[4];
  int nbufs = 0, nevs = 0;
  ~PipelineGuard() {
    for (int i = 0; i < nbufs; i++) scratch_free(bufs[i]);
    for (int i = 0; i < nevs; i++) cudaEventDestroy(evs[i]);
  }
  int alloc(void** p, size_t bytes) {
    const int rc = scratch_alloc(p, bytes);
    if (rc == TACO_B200_OK && nbufs < 32) bufs[nbufs++] = *p;
    return rc;
  }
  int event(cudaEvent_t* e) {
    if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) return fail(TACO_B200_ERR_CUDA, "cudaEventCreate failed");
    if (nevs < 4) evs[nevs++] = *e;
    return TACO_B200_OK;
  }
};

// result allocation in the configured result space (HOST: malloc, DEVICE: cudaMalloc)
void* result_alloc(size_t bytes);
// pooled device arrays that may be handed to the caller (released with taco_b200_free / device_result_free)
int device_result_alloc(void** p, size_t bytes);
void device_result_free(void* p);
// copy a small device array to host synchronously (e.g. pos[n] after the scan)
int read_back(void* host, const void* dev, size_t bytes);
// device array -> freshly malloc'ed host array through pinned chunks + parallel host copies (synchronous)
int d2h_fresh(void* dst, const void* src, size_t bytes);
// read one int32 that may live on host or device
int read_i32(const int32_t* p, int32_t* out);

// ---------------------------------------------------------------------------------------------------------
// tensor views (abi.cu) -- validate a taco_tensor_t against the format a kernel family expects
// ---------------------------------------------------------------------------------------------------------
enum class DType { F32, F64 };
int dtype_of(const taco_tensor_t* t, DType* out);
inline size_t dsize(DType d) { return d == DType::F32 ? 4 : 8; }

struct DenseView {   // {Dense,...}: vals only
  int32_t order; int32_t dim[3]; int32_t mode_order[3]; void* vals; DType dt;
  size_t count() const { size_t n = 1; for (int i = 0; i < order; i++) n *= (size_t)dim[i]; return n; }
};
struct CsrView {     // {Dense, Compressed}, mode ordering {0,1}
  int32_t rows, cols; int32_t* pos; int32_t* crd; void* vals; DType dt;
};
struct Csf3View {    // {Compressed, Compressed, Compressed}, mode ordering {0,1,2}
  int32_t dim[3]; int32_t* pos[3]; int32_t* crd[3]; void* vals; DType dt;
};
int view_dense(const taco_tensor_t* t, int order, const char* name, DenseView* v);
int view_csr(const taco_tensor_t* t, const char* name, CsrView* v);
int view_csf3(const taco_tensor_t* t, const char* name, Csf3View* v);

}  // namespace tb

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
namespace tbd {

// streaming 128-bit loads that bypass L1 allocation (crd / vals are touched exactly once)
__device__ __forceinline__ int4 ldg_stream_i4(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ldg_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg_stream_d2(const void* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int ldg_stream_i32(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
// L2 cache policies (createpolicy): evict_last for the one operand with re-use, evict_first for streams
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// single-element streaming loads: no L1 allocation, first to leave L2
__device__ __forceinline__ int ldg_stream_i32(const int* p, uint64_t strm) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(strm));
  return r;
}
__device__ __forceinline__ float ldg_stream(const float* p, uint64_t strm) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(strm));
  return r;
}
__device__ __forceinline__ double ldg_stream(const double* p, uint64_t strm) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(strm));
  return r;
}
// streaming stores (results are written once, never re-read by the kernel)
__device__ __forceinline__ void stg_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_d2(void* p, double2 v) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// largest r in [lo, hi] with a[r] <= target   (a non-decreasing, a[lo] <= target assumed)
__device__ __forceinline__ int search_last_le(const int* __restrict__ a, int lo, int hi, int target) {
  while (lo < hi) {
    int mid = lo + ((hi - lo + 1) >> 1);
    if (__ldg(a + mid) <= target) lo = mid; else hi = mid - 1;
  }
  return lo;
}
// smallest r in [lo, hi] with a[r] >= target; returns hi+1 if none
__device__ __forceinline__ int search_first_ge(const int* __restrict__ a, int lo, int hi, int target) {
  int end = hi + 1;
  while (lo < end) {
    int mid = lo + ((end - lo) >> 1);
    if (__ldg(a + mid) >= target) end = mid; else lo = mid + 1;
  }
  return lo;
}

}  // namespace tbd
#endif
