// runtime.cu -- device-resident storage / launch layer of libtaco_b200.
//
// Replaces the reference's CUDA runtime shim (/root/reference/src/cuda.cpp:11-72: two global switches plus
// cudaMallocManaged/cudaFree) and the unified-memory allocation convention of
// /root/reference/src/taco_tensor_t.cpp:10-67 and src/storage/array.cpp:212-219.  Instead of managed memory that
// page-faults to the GPU on first touch, operands are explicit HBM residents:
//   * device pointers inside a taco_tensor_t are used in place,
//   * pinned host arrays are DMA'd on the compute stream,
//   * pageable host arrays go through cudaMemcpyAsync staging,
//   * arrays registered with taco_b200_make_resident() are uploaded once and reused across calls,
// and all temporaries come from the stream-ordered pool (cudaMallocAsync), so assemble/compute are stream-ordered
// end to end (the reference synchronises the whole device after every launch, codegen_cuda.cpp:608-609).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace tb {

static thread_local char g_err[1024] = "";
static thread_local bool g_need_sync = false;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct Resident { void* dptr; size_t bytes; };
struct ProfEntry {
  std::string name;
  double ms = 0;
  long launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

struct Context {
  std::mutex mu;
  bool ready = false;
  int device = 0;
  int sms = 148;
  size_t mem_total = (size_t)180 << 30;
  cudaStream_t own_stream = nullptr;
  cudaStream_t cur_stream = nullptr;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};   // upload / download streams of the host-operand pipelines, side compute stream
  int space = TACO_B200_SPACE_HOST;
  long launches = 0;
  std::unordered_map<const void*, Resident> resident;
  std::unordered_set<const void*> pooled;   // device result arrays handed to the caller that came from the pool
  bool profile = false;
  std::vector<ProfEntry> prof;
  const char* mc_local = nullptr;            // result fan-out window: local base, size, and either the multicast base
  size_t mc_bytes = 0;                       // (mc_peers == -1) or the bases of the same window on mc_peers peer GPUs
  int mc_peers = 0;
  char* mc_base[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
static Context g;

static int init_locked(int device) {
  if (g.ready) return TACO_B200_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(TACO_B200_ERR_CUDA, "no CUDA device available (%s); libtaco_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(TACO_B200_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
  TB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  TB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(TACO_B200_ERR_CUDA, "device %d is sm_%d%d; libtaco_b200 is built for sm_100a only", device, prop.major,
                prop.minor);
  g.device = device;
  g.sms = prop.multiProcessorCount;
  g.mem_total = prop.totalGlobalMem;
  TB_CUDA(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
  g.cur_stream = g.own_stream;
  cudaMemPool_t pool;
  TB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = UINT64_MAX;   // never trim the pool between calls: temporaries are recycled
  TB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  g.ready = true;
  return TACO_B200_OK;
}

int ensure_init() {
  if (g.ready) {
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != g.device) cudaSetDevice(g.device);
    return TACO_B200_OK;
  }
  std::lock_guard<std::mutex> lk(g.mu);
  int dev = 0;
  const char* env = getenv("TACO_B200_DEVICE");
  if (env) dev = atoi(env);
  else if ((env = getenv("LOCAL_RANK"))) {   // one process per GPU under torchrun
    int n = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) dev = atoi(env) % n;
  }
  return init_locked(dev);
}

cudaStream_t stream() { return g.cur_stream; }
cudaStream_t aux_stream(int i) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.aux[i]) cudaStreamCreateWithFlags(&g.aux[i], cudaStreamNonBlocking);
  return g.aux[i];
}
bool is_resident(const void* host_ptr, size_t bytes) {
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.resident.find(host_ptr);
  return it != g.resident.end() && it->second.bytes >= bytes;
}
int num_sms() { return g.sms; }
size_t device_mem_total() { return g.mem_total; }
void count_launch(int n) { g.launches += n; }
int result_space() { return g.space; }
bool need_sync() { return g_need_sync; }
void clear_need_sync() { g_need_sync = false; }

int finish_call() {
  if (g_need_sync) {
    g_need_sync = false;
    TB_CUDA(cudaStreamSynchronize(g.cur_stream));
  } else {
    TB_CUDA(cudaPeekAtLastError());
  }
  return TACO_B200_OK;
}

Mem classify(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return Mem::Host;
  }
  switch (a.type) {
    case cudaMemoryTypeDevice: return Mem::Device;
    case cudaMemoryTypeManaged: return Mem::Device;   // managed memory is dereferenceable on the device
    case cudaMemoryTypeHost: return Mem::Pinned;
    default: return Mem::Host;
  }
}

Fanout result_fanout(const void* p, size_t bytes) {
  Fanout fo;
  if (g.mc_peers == 0 || !p) return fo;
  const char* c = (const char*)p;
  if (c < g.mc_local || c + bytes > g.mc_local + g.mc_bytes) return fo;
  fo.n = g.mc_peers;
  for (int i = 0; i < (g.mc_peers < 0 ? 1 : g.mc_peers); i++) fo.d[i] = (long long)(g.mc_base[i] - g.mc_local);
  return fo;
}

bool trusts_vals_size(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice;
}

ProfScope::ProfScope(const char* kernel_name) {
  if (!g.profile) return;
  std::lock_guard<std::mutex> lk(g.mu);
  for (size_t i = 0; i < g.prof.size(); i++)
    if (g.prof[i].name == kernel_name) slot = (int)i;
  if (slot < 0) {
    g.prof.push_back(ProfEntry());
    g.prof.back().name = kernel_name;
    slot = (int)g.prof.size() - 1;
  }
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, g.cur_stream);
  g.prof[slot].pending.push_back({a, b});
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g.mu);
  cudaEventRecord(g.prof[slot].pending.back().second, g.cur_stream);
}

void* small_scratch() {
  static void* buf = nullptr;
  static std::once_flag once;
  std::call_once(once, [] { if (cudaMalloc(&buf, 256) != cudaSuccess) { buf = nullptr; cudaGetLastError(); } });
  return buf;
}
std::mutex& small_scratch_mutex() {
  static std::mutex mu;
  return mu;
}

int scratch_alloc(void** p, size_t bytes) {
  *p = nullptr;
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocAsync(p, bytes, g.cur_stream);
  if (e != cudaSuccess) return fail(TACO_B200_ERR_ALLOC, "cudaMallocAsync(%zu): %s", bytes, cudaGetErrorString(e));
  return TACO_B200_OK;
}
void scratch_free(void* p) {
  if (p) cudaFreeAsync(p, g.cur_stream);
}

int d2h_fresh(void* dst, const void* src, size_t bytes);
int h2d_pageable(void* dst, const void* src, size_t bytes);
In::~In() { scratch_free(owned); }
int In::acquire(const void* p, size_t bytes) {
  if (p == nullptr) return fail(TACO_B200_ERR_ARG, "NULL operand array");
  Mem m = classify(p);
  if (m == Mem::Device) { dptr = p; return TACO_B200_OK; }
  {
    std::lock_guard<std::mutex> lk(g.mu);
    auto it = g.resident.find(p);
    if (it != g.resident.end() && it->second.bytes >= bytes) { dptr = it->second.dptr; return TACO_B200_OK; }
  }
  TB_TRY(scratch_alloc(&owned, bytes));
  dptr = owned;
  if (bytes >= (8u << 20) && m == Mem::Host) return h2d_pageable(owned, p, bytes);     // pageable operand (taco's malloc'ed arrays)
  if (bytes) TB_CUDA(cudaMemcpyAsync(owned, p, bytes, cudaMemcpyHostToDevice, g.cur_stream));
  // a copy out of PINNED memory is truly asynchronous: the call must not return before the DMA has read the operand, or a
  // caller that updates x / B for its next step would race with it (pageable sources are staged before the call returns)
  if (bytes && m == Mem::Pinned) g_need_sync = true;
  return TACO_B200_OK;
}

Out::~Out() { scratch_free(owned); }
int Out::acquire(void* p, size_t nbytes) {
  if (p == nullptr) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
  bytes = nbytes;
  if (classify(p) == Mem::Device) { dptr = p; return TACO_B200_OK; }
  TB_TRY(scratch_alloc(&owned, nbytes));
  dptr = owned;
  host_dst = p;
  return TACO_B200_OK;
}
int Out::commit() {
  if (host_dst && bytes) {
    // pageable destination (taco's own malloc'ed result arrays): pinned staging + parallel host copies instead of the
    // driver's single-threaded staged copy (and its page faults, if the array was never touched)
    if (bytes >= (4u << 20) && classify(host_dst) == Mem::Host) return d2h_fresh(host_dst, dptr, bytes);
    TB_CUDA(cudaMemcpyAsync(host_dst, dptr, bytes, cudaMemcpyDeviceToHost, g.cur_stream));
    g_need_sync = true;
  }
  return TACO_B200_OK;
}

// Device array -> a freshly malloc'ed host array (the arrays a HOST-space assemble / pack hands to taco, which must be legal for
// free()).  A plain cudaMemcpyAsync into fresh pageable memory pays for the driver's staged copy AND one page fault per 4 KB on a
// single thread (measured: 124 MB in 26-60 ms).  Here the DMA lands in two persistent pinned chunks and a few host threads copy
// chunk c-1 into the destination (taking its page faults in parallel) while chunk c is in flight.
int d2h_fresh(void* dst, const void* src, size_t bytes) {
  if (bytes == 0) return TACO_B200_OK;
  constexpr size_t CH = 16u << 20;
  if (bytes < (4u << 20)) {
    TB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g.cur_stream));
    TB_CUDA(cudaStreamSynchronize(g.cur_stream));
    return TACO_B200_OK;
  }
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  static char* stage[2] = {nullptr, nullptr};
  static cudaEvent_t ev[2];
  if (!stage[0]) {
    for (int b = 0; b < 2; b++) {
      TB_CUDA(cudaHostAlloc((void**)&stage[b], CH, cudaHostAllocDefault));
      TB_CUDA(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
    }
  }
  const unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = hw >= 8 ? 8 : (hw > 1 ? (int)hw : 1);
  const size_t nchunks = (bytes + CH - 1) / CH;
  auto drain = [&](size_t c) -> int {           // chunk c has landed in stage[c & 1]: copy it out with nthreads threads
    const int b = (int)(c & 1);
    TB_CUDA(cudaEventSynchronize(ev[b]));
    const size_t off = c * CH, len = bytes - off < CH ? bytes - off : CH;
    const size_t per = ((len + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) {
      const size_t a = (size_t)t * per;
      if (a >= len) break;
      const size_t n = len - a < per ? len - a : per;
      th.emplace_back([=] { memcpy((char*)dst + off + a, stage[b] + a, n); });
    }
    memcpy((char*)dst + off, stage[b], len < per ? len : per);
    for (auto& x : th) x.join();
    return TACO_B200_OK;
  };
  for (size_t c = 0; c < nchunks; c++) {
    const int b = (int)(c & 1);
    const size_t off = c * CH, len = bytes - off < CH ? bytes - off : CH;
    TB_CUDA(cudaMemcpyAsync(stage[b], (const char*)src + off, len, cudaMemcpyDeviceToHost, g.cur_stream));
    TB_CUDA(cudaEventRecord(ev[b], g.cur_stream));
    if (c >= 1) TB_TRY(drain(c - 1));           // stage[b ^ 1] is free again before chunk c + 1 is issued into it
  }
  return drain(nchunks - 1);
}

// Pageable host array -> device: a few host threads copy chunk c into one of two pinned buffers while the DMA of chunk c-1 runs
// (the driver's own staged copy is single-threaded: ~14 GB/s measured against 55 GB/s from pinned memory).
int h2d_pageable(void* dst, const void* src, size_t bytes) {
  constexpr size_t CH = 16u << 20;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  static char* stage[2] = {nullptr, nullptr};
  static cudaEvent_t ev[2];
  if (!stage[0]) {
    for (int b = 0; b < 2; b++) {
      TB_CUDA(cudaHostAlloc((void**)&stage[b], CH, cudaHostAllocDefault));
      TB_CUDA(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
      TB_CUDA(cudaEventRecord(ev[b], g.cur_stream));
    }
  }
  const unsigned hw = std::thread::hardware_concurrency();
  const int nthreads = hw >= 8 ? 8 : (hw > 1 ? (int)hw : 1);
  const size_t nchunks = (bytes + CH - 1) / CH;
  for (size_t c = 0; c < nchunks; c++) {
    const int b = (int)(c & 1);
    const size_t off = c * CH, len = bytes - off < CH ? bytes - off : CH;
    TB_CUDA(cudaEventSynchronize(ev[b]));       // the DMA that last read stage[b] has finished
    const size_t per = ((len + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) {
      const size_t a = (size_t)t * per;
      if (a >= len) break;
      const size_t n = len - a < per ? len - a : per;
      th.emplace_back([=] { memcpy(stage[b] + a, (const char*)src + off + a, n); });
    }
    memcpy(stage[b], (const char*)src + off, len < per ? len : per);
    for (auto& x : th) x.join();
    TB_CUDA(cudaMemcpyAsync((char*)dst + off, stage[b], len, cudaMemcpyHostToDevice, g.cur_stream));
    TB_CUDA(cudaEventRecord(ev[b], g.cur_stream));
  }
  return TACO_B200_OK;
}

// Device result arrays come from the stream-ordered pool (release threshold = never trim), so a loop of
// assemble/compute/free calls recycles the same HBM without cudaMalloc/cudaFree (which synchronise the device).
int device_result_alloc(void** p, size_t bytes) {
  TB_TRY(scratch_alloc(p, bytes));
  std::lock_guard<std::mutex> lk(g.mu);
  g.pooled.insert(*p);
  return TACO_B200_OK;
}
void device_result_free(void* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g.mu);
    g.pooled.erase(p);
  }
  cudaFreeAsync(p, g.cur_stream);
}

void* result_alloc(size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (g.space == TACO_B200_SPACE_DEVICE) {
    void* p = nullptr;
    if (device_result_alloc(&p, bytes) != TACO_B200_OK) { cudaGetLastError(); return nullptr; }
    return p;
  }
  return malloc(bytes);
}

int read_back(void* host, const void* dev, size_t bytes) {
  TB_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, g.cur_stream));
  TB_CUDA(cudaStreamSynchronize(g.cur_stream));
  return TACO_B200_OK;
}

int read_i32(const int32_t* p, int32_t* out) {
  if (classify(p) == Mem::Device) return read_back(out, p, sizeof(int32_t));
  *out = *p;
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

const char* taco_b200_last_error(void) { return g_err; }
const char* taco_b200_version(void) { return "taco_b200 0.1 (sm_100a)"; }

int taco_b200_init(int device) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.ready && device != g.device)
    return fail(TACO_B200_ERR_ARG, "already initialised on device %d (one process per GPU)", g.device);
  return init_locked(device);
}

int taco_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int taco_b200_set_stream(void* s) {
  TB_TRY(ensure_init());
  g.cur_stream = s ? (cudaStream_t)s : g.own_stream;
  return TACO_B200_OK;
}
void* taco_b200_get_stream(void) { return ensure_init() == TACO_B200_OK ? (void*)g.cur_stream : nullptr; }

int taco_b200_synchronize(void) {
  TB_TRY(ensure_init());
  TB_CUDA(cudaStreamSynchronize(g.cur_stream));
  return TACO_B200_OK;
}

int taco_b200_launch_count(void) { return (int)g.launches; }

int taco_b200_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g.mu);
  g.profile = on != 0;
  return TACO_B200_OK;
}

static void prof_drain(ProfEntry& e) {
  for (auto& pr : e.pending) {
    float ms = 0;
    if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
      e.ms += ms;
      e.launches++;
    }
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  e.pending.clear();
}

int taco_b200_profile_get(const char* kernel_name, double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(g.mu);
  for (auto& e : g.prof) {
    if (e.name == kernel_name) {
      prof_drain(e);
      if (total_ms) *total_ms = e.ms;
      if (launches) *launches = (int)e.launches;
      return TACO_B200_OK;
    }
  }
  return fail(TACO_B200_ERR_ARG, "no profile entry named '%s'", kernel_name);
}

int taco_b200_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  for (auto& e : g.prof) prof_drain(e);
  g.prof.clear();
  return TACO_B200_OK;
}

int taco_b200_set_result_space(int space) {
  if (space != TACO_B200_SPACE_HOST && space != TACO_B200_SPACE_DEVICE)
    return fail(TACO_B200_ERR_ARG, "unknown result space %d", space);
  g.space = space;
  return TACO_B200_OK;
}
int taco_b200_get_result_space(void) { return g.space; }

void* taco_b200_host_alloc(size_t bytes) {
  if (ensure_init() != TACO_B200_OK) return nullptr;
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) {
    fail(TACO_B200_ERR_ALLOC, "cudaHostAlloc(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void taco_b200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
void* taco_b200_device_alloc(size_t bytes) {
  if (ensure_init() != TACO_B200_OK) return nullptr;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
    fail(TACO_B200_ERR_ALLOC, "cudaMalloc(%zu) failed", bytes);
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void taco_b200_free(void* p) {
  if (!p) return;
  {
    std::unique_lock<std::mutex> lk(g.mu);
    if (g.pooled.count(p)) {
      lk.unlock();
      device_result_free(p);      // stream-ordered: returns to the pool after the work already enqueued
      return;
    }
  }
  Mem m = classify(p);
  if (m == Mem::Device) cudaFree(p);
  else if (m == Mem::Pinned) cudaFreeHost(p);
  else free(p);
}

int taco_b200_set_result_multicast(const void* local_base, void* multicast_base, size_t bytes) {
  TB_TRY(ensure_init());
  if ((local_base == nullptr) != (multicast_base == nullptr)) return fail(TACO_B200_ERR_ARG, "set_result_multicast: both bases or neither");
  if (local_base && (((uintptr_t)local_base | (uintptr_t)multicast_base) & 15))
    return fail(TACO_B200_ERR_ARG, "set_result_multicast: bases must be 16-byte aligned");
  std::lock_guard<std::mutex> lk(g.mu);
  g.mc_local = (const char*)local_base;
  g.mc_base[0] = (char*)multicast_base;
  g.mc_bytes = local_base ? bytes : 0;
  g.mc_peers = local_base ? -1 : 0;
  return TACO_B200_OK;
}

int taco_b200_set_result_peers(const void* local_base, size_t bytes, int npeers, void* const* peer_bases) {
  TB_TRY(ensure_init());
  if (!local_base || npeers == 0) {
    std::lock_guard<std::mutex> lk(g.mu);
    g.mc_local = nullptr; g.mc_bytes = 0; g.mc_peers = 0;
    return TACO_B200_OK;
  }
  if (npeers < 0 || npeers > 7 || !peer_bases) return fail(TACO_B200_ERR_ARG, "set_result_peers: 1..7 peer windows");
  if ((uintptr_t)local_base & 15) return fail(TACO_B200_ERR_ARG, "set_result_peers: bases must be 16-byte aligned");
  for (int i = 0; i < npeers; i++)
    if (!peer_bases[i] || ((uintptr_t)peer_bases[i] & 15) || peer_bases[i] == local_base)
      return fail(TACO_B200_ERR_ARG, "set_result_peers: peer bases must be non-NULL, 16-byte aligned and differ from the local base");
  std::lock_guard<std::mutex> lk(g.mu);
  g.mc_local = (const char*)local_base;
  g.mc_bytes = bytes;
  g.mc_peers = npeers;
  for (int i = 0; i < npeers; i++) g.mc_base[i] = (char*)peer_bases[i];
  return TACO_B200_OK;
}

int taco_b200_make_resident(const void* host_ptr, size_t bytes) {
  TB_TRY(ensure_init());
  if (!host_ptr) return fail(TACO_B200_ERR_ARG, "NULL host_ptr");
  if (classify(host_ptr) == Mem::Device) return TACO_B200_OK;
  void* d = nullptr;
  TB_CUDA(cudaMalloc(&d, bytes ? bytes : 16));
  cudaError_t e = cudaMemcpyAsync(d, host_ptr, bytes, cudaMemcpyHostToDevice, g.cur_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(g.cur_stream);
  if (e != cudaSuccess) { cudaFree(d); return fail(TACO_B200_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e)); }
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.resident.find(host_ptr);
  if (it != g.resident.end()) cudaFree(it->second.dptr);
  g.resident[host_ptr] = Resident{d, bytes};
  return TACO_B200_OK;
}

int taco_b200_invalidate(const void* host_ptr) {
  std::lock_guard<std::mutex> lk(g.mu);
  auto it = g.resident.find(host_ptr);
  if (it != g.resident.end()) {
    cudaFree(it->second.dptr);
    g.resident.erase(it);
  }
  return TACO_B200_OK;
}

int taco_b200_drop_all_resident(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  for (auto& kv : g.resident) cudaFree(kv.second.dptr);
  g.resident.clear();
  return TACO_B200_OK;
}

}  // extern "C"
