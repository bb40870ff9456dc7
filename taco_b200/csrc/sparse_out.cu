// sparse_out.cu -- GPU assembly of sparse results: SpAdd  C(i,j) = A(i,j) + B(i,j)  and
//                                                   SpGEMM C(i,k) = A(i,j) * B(j,k),   all CSR.
//
// The reference has NO GPU path for sparse outputs: workspaces are disabled under CUDA
// (/root/reference/src/lower/lowerer_impl_imperative.cpp:2286-2290) and assembly is a host-serial append loop with
// realloc-by-copy on managed memory (src/codegen/codegen_cuda.cpp:1071-1135).  This file is the "src/lower gains GPU
// assembly" item of the north star: the two-phase Insert strategy the reference lowers for CPUs
// (lowerAssemble :2616-2779; CompressedModeFormat getSeqInitEdges/getSeqInsertEdge/getYieldPos/getFinalizeYieldPos,
// src/lower/mode_format_compressed.cpp:217-271) executed on the device:
//     symbolic   per-row size of the result pattern            (kernel, rows in parallel)
//     scan       pos = exclusive prefix sum of the sizes        (scan.cuh)
//     fill       crd written in ascending column order          (kernel)
//     numeric    values written at the positions pos/crd define (kernel; `compute`)
// Structure rules that make pos/crd BIT-EXACT with the reference: SpAdd = two-finger union keeping explicit zeros;
// SpGEMM = sorted set of reachable columns per row, entries that sum to zero are kept.
// SpGEMM values are accumulated in the reference's order (A-row order, then B-row order) -> bit-identical.
#include <climits>
#include <cstring>

#include "common.cuh"
#include "scan.cuh"

namespace tb {

int csr_nnz(const CsrView& A, int32_t vals_size_hint, int32_t* nnz);   // spmv.cu

// =========================================================================================================
// SpAdd
// =========================================================================================================
// One thread per row: both operand rows are short contiguous runs, so a warp touches one contiguous window of each
// crd/vals array (L1 absorbs the per-thread strides).
__global__ void __launch_bounds__(256)
spadd_count_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const int* __restrict__ Bpos,
                   const int* __restrict__ Bcrd, int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (i == n) { counts[n] = 0; return; }
  int a = __ldg(Apos + i), ae = __ldg(Apos + i + 1), b = __ldg(Bpos + i), be = __ldg(Bpos + i + 1), c = 0;
  if (a < ae && b < be) {
    int ja = __ldg(Acrd + a), jb = __ldg(Bcrd + b);
    while (true) {
      c++;
      int j = min(ja, jb);
      if (ja == j) { if (++a == ae) { if (jb == j) ++b; break; } ja = __ldg(Acrd + a); }
      if (jb == j) { if (++b == be) break; jb = __ldg(Bcrd + b); }
    }
  }
  counts[i] = c + (ae - a) + (be - b);
}

template <typename T, bool CRD, bool VALS>
__global__ void __launch_bounds__(256)
spadd_fill_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const T* __restrict__ Av,
                  const int* __restrict__ Bpos, const int* __restrict__ Bcrd, const T* __restrict__ Bv,
                  const int* __restrict__ Cpos, int* __restrict__ Ccrd, T* __restrict__ Cv) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = __ldg(Apos + i), ae = __ldg(Apos + i + 1), b = __ldg(Bpos + i), be = __ldg(Bpos + i + 1), p = __ldg(Cpos + i);
  while (a < ae && b < be) {
    int ja = __ldg(Acrd + a), jb = __ldg(Bcrd + b), j = min(ja, jb);
    if (CRD) Ccrd[p] = j;
    if (VALS) Cv[p] = (ja == j && jb == j) ? (__ldg(Av + a) + __ldg(Bv + b)) : (ja == j ? __ldg(Av + a) : __ldg(Bv + b));
    p++;
    a += (ja == j);
    b += (jb == j);
  }
  for (; a < ae; a++, p++) { if (CRD) Ccrd[p] = __ldg(Acrd + a); if (VALS) Cv[p] = __ldg(Av + a); }
  for (; b < be; b++, p++) { if (CRD) Ccrd[p] = __ldg(Bcrd + b); if (VALS) Cv[p] = __ldg(Bv + b); }
}

// =========================================================================================================
// SpGEMM
// =========================================================================================================
constexpr int SG_WARP_CAP = 256;     // products per row handled by a warp team in shared memory
constexpr int SG_CTA_CAP = 8192;     // products per row handled by a CTA team in shared memory

// upper bound of the row pattern = number of products; rows are binned by it
__global__ void __launch_bounds__(256)
spgemm_bound_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const int* __restrict__ Bpos,
                    int* __restrict__ bin_count, int* __restrict__ bin_rows, int* __restrict__ counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long ub = 0;
  for (int p = __ldg(Apos + i); p < __ldg(Apos + i + 1); p++) {
    int j = __ldg(Acrd + p);
    ub += __ldg(Bpos + j + 1) - __ldg(Bpos + j);
  }
  if (ub == 0) { counts[i] = 0; return; }
  int bin = ub <= SG_WARP_CAP ? 0 : (ub <= SG_CTA_CAP ? 1 : 2);
  int slot = atomicAdd(bin_count + bin, 1);
  bin_rows[(size_t)bin * n + slot] = i;
}

// Team-cooperative symbolic phase: expand the columns of all products of a row into shared memory, bitonic-sort them,
// count (WRITE=false) or emit (WRITE=true) the distinct ones in ascending order.
template <int THREADS, int CAP, bool WRITE>
__global__ void __launch_bounds__(256)
spgemm_symbolic_kernel(const int* __restrict__ rows_list, int nrows_bin, const int* __restrict__ Apos,
                       const int* __restrict__ Acrd, const int* __restrict__ Bpos, const int* __restrict__ Bcrd,
                       int* __restrict__ counts, const int* __restrict__ Cpos, int* __restrict__ Ccrd) {
  constexpr int TEAMS = 256 / THREADS;
  constexpr int CHUNK = CAP / THREADS;              // keys per thread in the compaction phase
  __shared__ int keys_all[TEAMS][CAP];
  __shared__ int aux_all[TEAMS][THREADS + 1];
  const int team = threadIdx.x / THREADS, tid = threadIdx.x % THREADS;
  const int ridx = blockIdx.x * TEAMS + team;
  auto team_sync = [&]() { if (THREADS == 32) __syncwarp(); else __syncthreads(); };
  // (a CTA team never exits early: TEAMS == 1 and the grid is exact; warp teams may return as a whole warp)
  if (ridx >= nrows_bin) return;
  int* keys = keys_all[team];
  int* aux = aux_all[team];
  const int i = __ldg(rows_list + ridx);
  const int a0 = __ldg(Apos + i), a1 = __ldg(Apos + i + 1);
  // ---- expand: THREADS A-entries at a time; a team-wide scan of the B-row lengths gives the write offsets -------
  int total = 0;
  for (int ab = a0; ab < a1; ab += THREADS) {
    int bs = 0, len = 0;
    if (ab + tid < a1) {
      int j = __ldg(Acrd + ab + tid);
      bs = __ldg(Bpos + j);
      len = __ldg(Bpos + j + 1) - bs;
    }
    aux[tid + 1] = len;
    team_sync();
    if (tid == 0) {
      aux[0] = 0;
      for (int t = 1; t <= THREADS; t++) aux[t] += aux[t - 1];
    }
    team_sync();
    const int off = total + aux[tid];
    // each thread copies its own B row: rows are short at the target shapes; long rows still coalesce per thread
    for (int t = 0; t < len; t++) keys[off + t] = __ldg(Bcrd + bs + t);
    total += aux[THREADS];
    team_sync();
  }
  int m = 32;
  while (m < total) m <<= 1;
  for (int t = total + tid; t < m; t += THREADS) keys[t] = INT_MAX;
  team_sync();
  // ---- bitonic sort of m keys -----------------------------------------------------------------------------
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < (m >> 1); t += THREADS) {
        int lo = 2 * j * (t / j) + (t % j), hi = lo + j;
        int x = keys[lo], y = keys[hi];
        bool up = (lo & k) == 0;
        if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
      }
      team_sync();
    }
  }
  // ---- distinct keys: contiguous chunk per thread, team scan of the chunk counts -------------------------------
  const int c0 = tid * CHUNK;
  int cnt = 0;
  for (int t = c0; t < c0 + CHUNK && t < total; t++) cnt += (t == 0 || keys[t] != keys[t - 1]);
  if (!WRITE) {
    aux[tid] = cnt;
    team_sync();
    if (tid == 0) {
      int s = 0;
      for (int t = 0; t < THREADS; t++) s += aux[t];
      counts[i] = s;
    }
  } else {
    aux[tid + 1] = cnt;
    team_sync();
    if (tid == 0) {
      aux[0] = 0;
      for (int t = 1; t <= THREADS; t++) aux[t] += aux[t - 1];
    }
    team_sync();
    int* out = Ccrd + __ldg(Cpos + i) + aux[tid];
    for (int t = c0; t < c0 + CHUNK && t < total; t++)
      if (t == 0 || keys[t] != keys[t - 1]) *out++ = keys[t];
  }
}

// Rows with more than SG_CTA_CAP products: a bitmap over the column space in global scratch (one per CTA) gives the
// sorted distinct columns directly -- the device form of the reference's w_already_set / w_index_list + qsort.
template <bool WRITE>
__global__ void __launch_bounds__(256)
spgemm_symbolic_bitmap_kernel(const int* __restrict__ rows_list, int nrows_bin, int ncols, const int* __restrict__ Apos,
                              const int* __restrict__ Acrd, const int* __restrict__ Bpos, const int* __restrict__ Bcrd,
                              unsigned* __restrict__ bitmaps, int* __restrict__ counts, const int* __restrict__ Cpos,
                              int* __restrict__ Ccrd) {
  __shared__ int wsum[8];
  __shared__ int running;
  const int words = (ncols + 31) / 32;
  unsigned* bm = bitmaps + (size_t)blockIdx.x * words;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int ridx = blockIdx.x; ridx < nrows_bin; ridx += gridDim.x) {
    const int i = __ldg(rows_list + ridx);
    for (int t = threadIdx.x; t < words; t += 256) bm[t] = 0u;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (int pa = __ldg(Apos + i) + wid; pa < __ldg(Apos + i + 1); pa += 8) {   // one warp per A entry
      const int j = __ldg(Acrd + pa);
      for (int pb = __ldg(Bpos + j) + lane; pb < __ldg(Bpos + j + 1); pb += 32) {
        int k = __ldg(Bcrd + pb);
        atomicOr(bm + (k >> 5), 1u << (k & 31));
      }
    }
    __syncthreads();
    for (int wb = 0; wb < words; wb += 256) {
      const int t = wb + threadIdx.x;
      const unsigned bits = t < words ? bm[t] : 0u;
      const int c = __popc(bits);
      int incl = c;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
      }
      if (lane == 31) wsum[wid] = incl;
      __syncthreads();
      int before = running;
      for (int q = 0; q < wid; q++) before += wsum[q];
      if (WRITE) {
        int* out = Ccrd + __ldg(Cpos + i) + before + incl - c;
        unsigned b = bits;
        while (b) { int bit = __ffs(b) - 1; *out++ = t * 32 + bit; b &= b - 1; }
      }
      __syncthreads();
      if (threadIdx.x == 255) running = before + incl;
      __syncthreads();
    }
    if (!WRITE && threadIdx.x == 0) counts[i] = running;
    __syncthreads();
  }
}

// Numeric phase.  A team of TEAM lanes owns a row.  A-entries are taken one after the other (the reference's outer
// order); the lanes of the team take the entries of that B row (distinct columns => no write conflicts), locate the
// column in the row's sorted crd by binary search and accumulate into C_vals -- first touch stores a*b, later touches
// add, exactly as the reference's workspace does (w[k] = a*b | w[k] = w[k] + a*b).
template <typename T, int TEAM>
__global__ void __launch_bounds__(256)
spgemm_numeric_kernel(int n, const int* __restrict__ Apos, const int* __restrict__ Acrd, const T* __restrict__ Av,
                      const int* __restrict__ Bpos, const int* __restrict__ Bcrd, const T* __restrict__ Bv,
                      const int* __restrict__ Cpos, const int* __restrict__ Ccrd, T* Cv) {
  const int gt = blockIdx.x * 256 + threadIdx.x;
  const int i = gt / TEAM, tl = gt % TEAM;
  const unsigned mask = TEAM == 32 ? 0xffffffffu : (((1u << TEAM) - 1u) << ((threadIdx.x & 31) / TEAM * TEAM));
  if (i >= n) return;
  const int c0 = __ldg(Cpos + i), c1 = __ldg(Cpos + i + 1);
  for (int t = c0 + tl; t < c1; t += TEAM) Cv[t] = T(0);
  __syncwarp(mask);
  const int a0 = __ldg(Apos + i), a1 = __ldg(Apos + i + 1);
  for (int ab = a0; ab < a1; ab += TEAM) {
    int my_bs = 0, my_be = 0;
    T my_a = T(0);
    if (ab + tl < a1) {
      int j = __ldg(Acrd + ab + tl);
      my_a = __ldg(Av + ab + tl);
      my_bs = __ldg(Bpos + j);
      my_be = __ldg(Bpos + j + 1);
    }
    const int cnt = min(TEAM, a1 - ab);
    for (int q = 0; q < cnt; q++) {
      const int src = (threadIdx.x & 31) / TEAM * TEAM + q;
      const int bs = __shfl_sync(mask, my_bs, src), be = __shfl_sync(mask, my_be, src);
      const T a = __shfl_sync(mask, my_a, src);
      for (int pb = bs + tl; pb < be; pb += TEAM) {
        const int k = __ldg(Bcrd + pb);
        int lo = c0, hi = c1 - 1;                  // k is guaranteed present in Ccrd[c0..c1)
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (__ldg(Ccrd + mid) < k) lo = mid + 1; else hi = mid;
        }
        Cv[lo] = Cv[lo] + a * __ldg(Bv + pb);
      }
      __syncwarp(mask);
    }
  }
}

// =========================================================================================================
// host side
// =========================================================================================================
struct Csr3 { CsrView C, A, B; int32_t nnzA, nnzB; In apos, acrd, bpos, bcrd; };

static int csr3_prepare(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, bool product, Csr3* s) {
  TB_TRY(ensure_init());
  TB_TRY(view_csr(C, "C", &s->C));
  TB_TRY(view_csr(A, "A", &s->A));
  TB_TRY(view_csr(B, "B", &s->B));
  if (product) {
    if (s->A.cols != s->B.rows || s->C.rows != s->A.rows || s->C.cols != s->B.cols)
      return fail(TACO_B200_ERR_ARG, "spgemm: dimension mismatch");
  } else if (s->A.rows != s->B.rows || s->A.cols != s->B.cols || s->C.rows != s->A.rows || s->C.cols != s->A.cols) {
    return fail(TACO_B200_ERR_ARG, "spadd: dimension mismatch");
  }
  if (s->A.dt != s->B.dt || s->C.dt != s->A.dt) return fail(TACO_B200_ERR_FORMAT, "mixed component types");
  TB_TRY(csr_nnz(s->A, A->vals_size, &s->nnzA));
  TB_TRY(csr_nnz(s->B, B->vals_size, &s->nnzB));
  TB_TRY(s->apos.acquire(s->A.pos, sizeof(int32_t) * ((size_t)s->A.rows + 1)));
  TB_TRY(s->acrd.acquire(s->A.crd ? (void*)s->A.crd : (void*)s->A.pos, sizeof(int32_t) * (size_t)s->nnzA));
  TB_TRY(s->bpos.acquire(s->B.pos, sizeof(int32_t) * ((size_t)s->B.rows + 1)));
  TB_TRY(s->bcrd.acquire(s->B.crd ? (void*)s->B.crd : (void*)s->B.pos, sizeof(int32_t) * (size_t)s->nnzB));
  return TACO_B200_OK;
}

// hand freshly built device pos / crd to the caller in the configured result space
static int publish_structure(taco_tensor_t* C, int n, int* dpos, int* dcrd, int32_t nnzC, size_t esize) {
  if (result_space() == TACO_B200_SPACE_DEVICE) {
    void* vals = result_alloc(esize * (size_t)nnzC);
    if (!vals) return fail(TACO_B200_ERR_ALLOC, "cannot allocate result values");
    C->indices[1][0] = (uint8_t*)dpos;
    C->indices[1][1] = (uint8_t*)dcrd;
    C->vals = (uint8_t*)vals;
  } else {
    int32_t* hpos = (int32_t*)malloc(sizeof(int32_t) * ((size_t)n + 1));
    int32_t* hcrd = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnzC > 0 ? nnzC : 1));
    void* vals = malloc(esize * (size_t)(nnzC > 0 ? nnzC : 1));
    if (!hpos || !hcrd || !vals) return fail(TACO_B200_ERR_ALLOC, "cannot allocate host result arrays");
    TB_CUDA(cudaMemcpyAsync(hpos, dpos, sizeof(int32_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, stream()));
    if (nnzC) TB_CUDA(cudaMemcpyAsync(hcrd, dcrd, sizeof(int32_t) * (size_t)nnzC, cudaMemcpyDeviceToHost, stream()));
    TB_CUDA(cudaStreamSynchronize(stream()));
    device_result_free(dpos);
    device_result_free(dcrd);
    C->indices[1][0] = (uint8_t*)hpos;
    C->indices[1][1] = (uint8_t*)hcrd;
    C->vals = (uint8_t*)vals;
  }
  C->vals_size = nnzC;
  return TACO_B200_OK;
}

static int spadd_assemble_impl(taco_tensor_t* C, Csr3& s) {
  const int n = s.A.rows;
  int* dpos = nullptr;
  TB_TRY(device_result_alloc((void**)&dpos, sizeof(int) * ((size_t)n + 1)));
  {
    ProfScope ps("spadd_symbolic");
    spadd_count_kernel<<<(n + 1 + 255) / 256, 256, 0, stream()>>>(n, s.apos.as<int>(), s.acrd.as<int>(), s.bpos.as<int>(),
                                                                 s.bcrd.as<int>(), dpos);
  }
  count_launch(1);
  TB_TRY(exclusive_scan_i32(dpos, dpos, (long long)n + 1));
  int32_t nnzC = 0;
  TB_TRY(read_back(&nnzC, dpos + n, sizeof(int32_t)));
  int* dcrd = nullptr;
  TB_TRY(device_result_alloc((void**)&dcrd, sizeof(int) * (size_t)(nnzC > 0 ? nnzC : 1)));
  if (n > 0) {
    spadd_fill_kernel<double, true, false><<<(n + 255) / 256, 256, 0, stream()>>>(
        n, s.apos.as<int>(), s.acrd.as<int>(), nullptr, s.bpos.as<int>(), s.bcrd.as<int>(), nullptr, dpos, dcrd, nullptr);
    count_launch(1);
  }
  TB_CUDA(cudaGetLastError());
  return publish_structure(C, n, dpos, dcrd, nnzC, dsize(s.A.dt));
}

template <typename T>
static int spadd_numeric(Csr3& s, const In& av, const In& bv, const In& cpos, Out& cv) {
  const int n = s.A.rows;
  if (n > 0) {
    ProfScope ps("spadd_numeric");
    spadd_fill_kernel<T, false, true><<<(n + 255) / 256, 256, 0, stream()>>>(
        n, s.apos.as<int>(), s.acrd.as<int>(), av.as<T>(), s.bpos.as<int>(), s.bcrd.as<int>(), bv.as<T>(), cpos.as<int>(),
        nullptr, cv.as<T>());
    count_launch(1);
  }
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

static int spgemm_assemble_impl(taco_tensor_t* C, Csr3& s) {
  const int n = s.A.rows, ncols = s.B.cols;
  int* dpos = nullptr;
  TB_TRY(device_result_alloc((void**)&dpos, sizeof(int) * ((size_t)n + 1)));
  void *bin_count = nullptr, *bin_rows = nullptr, *bitmaps = nullptr;
  TB_TRY(scratch_alloc(&bin_count, sizeof(int) * 4));
  TB_TRY(scratch_alloc(&bin_rows, sizeof(int) * 3 * (size_t)(n > 0 ? n : 1)));
  TB_CUDA(cudaMemsetAsync(bin_count, 0, sizeof(int) * 4, stream()));
  TB_CUDA(cudaMemsetAsync(dpos + n, 0, sizeof(int), stream()));
  int hbin[4] = {0, 0, 0, 0};
  if (n > 0) {
    spgemm_bound_kernel<<<(n + 255) / 256, 256, 0, stream()>>>(n, s.apos.as<int>(), s.acrd.as<int>(), s.bpos.as<int>(),
                                                              (int*)bin_count, (int*)bin_rows, dpos);
    count_launch(1);
    TB_TRY(read_back(hbin, bin_count, sizeof(int) * 4));
  }
  const int* rows0 = (const int*)bin_rows;
  const int* rows1 = rows0 + n;
  const int* rows2 = rows1 + n;
  int big_grid = 0;
  if (hbin[2] > 0) {
    big_grid = hbin[2] < 2 * num_sms() ? hbin[2] : 2 * num_sms();
    TB_TRY(scratch_alloc(&bitmaps, sizeof(unsigned) * (size_t)big_grid * ((ncols + 31) / 32)));
  }
  for (int pass = 0; pass < 2; pass++) {
    int* dcrd = nullptr;
    int32_t nnzC = 0;
    if (pass == 1) {
      TB_TRY(exclusive_scan_i32(dpos, dpos, (long long)n + 1));
      TB_TRY(read_back(&nnzC, dpos + n, sizeof(int32_t)));
      TB_TRY(device_result_alloc((void**)&dcrd, sizeof(int) * (size_t)(nnzC > 0 ? nnzC : 1)));
    }
#define TB_SG_ARGS s.apos.as<int>(), s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), dpos, dpos, dcrd
    ProfScope ps("spgemm_symbolic");
    if (hbin[0] > 0) {
      int grid = (hbin[0] + 7) / 8;
      if (pass == 0) spgemm_symbolic_kernel<32, SG_WARP_CAP, false><<<grid, 256, 0, stream()>>>(rows0, hbin[0], TB_SG_ARGS);
      else spgemm_symbolic_kernel<32, SG_WARP_CAP, true><<<grid, 256, 0, stream()>>>(rows0, hbin[0], TB_SG_ARGS);
      count_launch(1);
    }
    if (hbin[1] > 0) {
      if (pass == 0) spgemm_symbolic_kernel<256, SG_CTA_CAP, false><<<hbin[1], 256, 0, stream()>>>(rows1, hbin[1], TB_SG_ARGS);
      else spgemm_symbolic_kernel<256, SG_CTA_CAP, true><<<hbin[1], 256, 0, stream()>>>(rows1, hbin[1], TB_SG_ARGS);
      count_launch(1);
    }
#undef TB_SG_ARGS
    if (hbin[2] > 0) {
      if (pass == 0)
        spgemm_symbolic_bitmap_kernel<false><<<big_grid, 256, 0, stream()>>>(rows2, hbin[2], ncols, s.apos.as<int>(),
            s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), (unsigned*)bitmaps, dpos, dpos, dcrd);
      else
        spgemm_symbolic_bitmap_kernel<true><<<big_grid, 256, 0, stream()>>>(rows2, hbin[2], ncols, s.apos.as<int>(),
            s.acrd.as<int>(), s.bpos.as<int>(), s.bcrd.as<int>(), (unsigned*)bitmaps, dpos, dpos, dcrd);
      count_launch(1);
    }
    TB_CUDA(cudaGetLastError());
    if (pass == 1) {
      scratch_free(bin_count); scratch_free(bin_rows); scratch_free(bitmaps);
      return publish_structure(C, n, dpos, dcrd, nnzC, dsize(s.A.dt));
    }
  }
  return TACO_B200_OK;
}

template <typename T>
static int spgemm_numeric(Csr3& s, const In& av, const In& bv, const In& cpos, const In& ccrd, Out& cv) {
  const int n = s.A.rows;
  if (n == 0) return TACO_B200_OK;
  double avg = s.B.rows > 0 ? (double)s.nnzB / s.B.rows : 0.0;
#define TB_SGN(TEAM)                                                                                               \
  spgemm_numeric_kernel<T, TEAM><<<(unsigned)(((long long)n * TEAM + 255) / 256), 256, 0, stream()>>>(              \
      n, s.apos.as<int>(), s.acrd.as<int>(), av.as<T>(), s.bpos.as<int>(), s.bcrd.as<int>(), bv.as<T>(), cpos.as<int>(), \
      ccrd.as<int>(), cv.as<T>())
  ProfScope ps("spgemm_numeric");
  if (avg <= 12.0) TB_SGN(8);
  else if (avg <= 24.0) TB_SGN(16);
  else TB_SGN(32);
#undef TB_SGN
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

static int sparse_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, bool product) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, product, &s));
  if (!s.C.pos) return fail(TACO_B200_ERR_ARG, "result has no structure: call assemble first");
  int32_t nnzC = 0;
  TB_TRY(csr_nnz(s.C, C->vals_size, &nnzC));
  size_t es = dsize(s.A.dt);
  In av, bv, cpos, ccrd; Out cv;
  TB_TRY(av.acquire(s.A.vals ? s.A.vals : (void*)s.A.pos, es * (size_t)s.nnzA));
  TB_TRY(bv.acquire(s.B.vals ? s.B.vals : (void*)s.B.pos, es * (size_t)s.nnzB));
  TB_TRY(cpos.acquire(s.C.pos, sizeof(int32_t) * ((size_t)s.C.rows + 1)));
  if (nnzC > 0) {
    TB_TRY(cv.acquire(s.C.vals, es * (size_t)nnzC));
    if (product) {
      TB_TRY(ccrd.acquire(s.C.crd, sizeof(int32_t) * (size_t)nnzC));
      if (s.A.dt == DType::F64) TB_TRY(spgemm_numeric<double>(s, av, bv, cpos, ccrd, cv));
      else TB_TRY(spgemm_numeric<float>(s, av, bv, cpos, ccrd, cv));
    } else {
      if (s.A.dt == DType::F64) TB_TRY(spadd_numeric<double>(s, av, bv, cpos, cv));
      else TB_TRY(spadd_numeric<float>(s, av, bv, cpos, cv));
    }
    TB_TRY(cv.commit());
  }
  return finish_call();
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_spadd_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, false, &s));
  return spadd_assemble_impl(C, s);
}
int taco_b200_spadd_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) { return sparse_compute(C, A, B, false); }
int taco_b200_spadd_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  TB_TRY(taco_b200_spadd_assemble(C, A, B));
  return taco_b200_spadd_compute(C, A, B);
}

int taco_b200_spgemm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  Csr3 s;
  TB_TRY(csr3_prepare(C, A, B, true, &s));
  return spgemm_assemble_impl(C, s);
}
int taco_b200_spgemm_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) { return sparse_compute(C, A, B, true); }
int taco_b200_spgemm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  TB_TRY(taco_b200_spgemm_assemble(C, A, B));
  return taco_b200_spgemm_compute(C, A, B);
}

}  // extern "C"
