// pack.cu -- COO -> the level arrays of a compressed format, on the device.
//
// Replaces TensorBase::pack() of the reference (/root/reference/src/tensor.cpp:295-463): a host qsort of the coordinate
// buffer (lexicographicalCmp) followed by the JIT-compiled `pack` helper (`int pack(taco_tensor_t* A, taco_tensor_t* B)`,
// src/tensor.cpp:932-1000, codegen.cpp:514-529), which walks the sorted coordinates, ADDS the values of equal
// coordinates and appends pos / crd / vals level by level.  The entry point here has the helper's signature and takes the
// same coordinate-buffer tensor (every level "sparse": indices[0][0] = {0, n}, indices[l][1] = the coordinates of level
// l, vals = the components), but the coordinates may arrive in ANY order:
//   1. keys: order 2 -> (c0 << b1) | c1; order 3 -> two stable passes, c2 first, then (c0 << b1) | c1.
//      Stable LSD radix sort over exactly the bits the dimensions need (cub::DeviceRadixSort -- library plumbing, like a
//      plain GEMM on cuBLAS; everything after it is this file's own kernels).
//   2. head flags per level (entry differs from its predecessor in the level's coordinate prefix), one exclusive scan
//      per level (scan.cuh) -> the node index of every entry at every level; the totals size the result arrays (one
//      small read-back, as the reference reads pos[parent] after assemble, src/tensor.cpp:263-292).
//   3. fill: crd of a level at the head entries, pos of a level at the heads of its parent level, values = the sum of a
//      run of equal coordinates in sorted (= insertion, the sort is stable) order.
// Targets: {Dense,Compressed} (CSR; CSC with mode ordering {1,0}), {Compressed,Compressed} (DCSR / DCSC), {Compressed x3}
// (CSF, any mode ordering): levels are taken in storage order, as the reference's helper receives them.
// Structure is bit-exact with the reference.  Values are bit-exact when coordinates are distinct; for duplicates the
// reference's summation order is that of an unstable qsort (unspecified), here it is insertion order.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace tb {

static int bits_for(int dim) {          // bits needed for coordinates 0 .. dim-1 (at least 1)
  int b = 1;
  while (b < 31 && (1ll << b) < (long long)dim) b++;
  return b;
}

__global__ void __launch_bounds__(256)
pack_keys_kernel(const int* __restrict__ c0, const int* __restrict__ c1, const unsigned* __restrict__ perm_in, int b1, long long n,
                 unsigned long long* __restrict__ key, unsigned* __restrict__ idx) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const unsigned src = perm_in ? perm_in[e] : (unsigned)e;
  key[e] = ((unsigned long long)(unsigned)__ldg(c0 + src) << b1) | (unsigned)__ldg(c1 + src);
  idx[e] = src;
}

__global__ void __launch_bounds__(256)
pack_key32_kernel(const int* __restrict__ c, long long n, unsigned* __restrict__ key, unsigned* __restrict__ idx) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  key[e] = (unsigned)__ldg(c + e);
  idx[e] = (unsigned)e;
}

// heads of the three possible levels for sorted entry e: bit 0 = new leaf (distinct coordinate), bit 1 = new (c0,c1)
// prefix (order 3 only: new fiber), bit 2 = new c0 (row / slice)
__global__ void __launch_bounds__(256)
pack_heads_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ c2, const unsigned* __restrict__ perm, int b1,
                  long long n, int* __restrict__ h_leaf, int* __restrict__ h_mid, int* __restrict__ h_top) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const unsigned long long k = key[e], kp = e ? key[e - 1] : ~0ull;
  const bool mid = e == 0 || k != kp;
  bool leaf = mid;
  if (c2 && !leaf) leaf = __ldg(c2 + perm[e]) != __ldg(c2 + perm[e - 1]);
  h_leaf[e] = leaf;
  if (h_mid) h_mid[e] = mid;
  h_top[e] = e == 0 || (k >> b1) != (kp >> b1);
}

struct PackOut {
  int* pos_top; int* crd_top;      // DCSR / CSF level 0 (pos_top has 2 entries) -- null for CSR
  int* pos_mid; int* crd_mid;      // CSF level 1 -- null for order 2
  int* pos_leaf; int* crd_leaf;    // last level
  int* run_start;                  // [nleaf + 1] first sorted entry of every distinct coordinate
};

// fill pass over the sorted entries (ids = exclusive scans of the head flags)
__global__ void __launch_bounds__(256)
pack_fill_kernel(const unsigned long long* __restrict__ key, const int* __restrict__ c2, const unsigned* __restrict__ perm, int b1,
                 long long n, const int* __restrict__ h_leaf, const int* __restrict__ h_mid, const int* __restrict__ h_top,
                 const int* __restrict__ id_leaf, const int* __restrict__ id_mid, const int* __restrict__ id_top, int rows_dense,
                 int nleaf, int nmid, int ntop, PackOut o) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const unsigned long long k = key[e];
  const int c0 = (int)(k >> b1), c1 = (int)(k & ((1ull << b1) - 1));
  const int u = id_leaf[e];
  if (h_leaf[e]) {
    o.run_start[u] = (int)e;
    o.crd_leaf[u] = c2 ? __ldg(c2 + perm[e]) : c1;
  }
  if (c2) {                                   // order 3: fibers (c0,c1) and slices c0
    if (h_mid[e]) { o.crd_mid[id_mid[e]] = c1; o.pos_leaf[id_mid[e]] = u; }
    if (h_top[e]) { o.crd_top[id_top[e]] = c0; o.pos_mid[id_top[e]] = id_mid[e]; }
  } else if (o.pos_top) {                     // DCSR: stored rows
    if (h_top[e]) { o.crd_top[id_top[e]] = c0; o.pos_leaf[id_top[e]] = u; }
  } else {                                    // CSR: dense rows; rows (previous row, c0] start at u
    if (h_top[e]) {
      const int prev = e ? (int)(key[e - 1] >> b1) : -1;
      for (int r = prev + 1; r <= c0; r++) o.pos_leaf[r] = u;
    }
    if (e == n - 1)
      for (int r = c0 + 1; r <= rows_dense; r++) o.pos_leaf[r] = nleaf;
  }
  if (e == n - 1) {                           // closing entries of the pos arrays
    o.run_start[nleaf] = (int)n;
    if (c2) { o.pos_leaf[nmid] = nleaf; o.pos_mid[ntop] = nmid; }
    else if (o.pos_top) o.pos_leaf[ntop] = nleaf;
    if (o.pos_top) { o.pos_top[0] = 0; o.pos_top[1] = ntop; }
  }
}

// one thread per distinct coordinate: the values of its run, added in sorted (insertion) order
template <typename T>
__global__ void __launch_bounds__(256)
pack_vals_kernel(const T* __restrict__ vals, const unsigned* __restrict__ perm, const int* __restrict__ run_start, int nleaf,
                 T* __restrict__ out) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nleaf) return;
  const int a = run_start[u], b = run_start[u + 1];
  T acc = __ldg(vals + perm[a]);
  for (int e = a + 1; e < b; e++) acc = acc + __ldg(vals + perm[e]);
  out[u] = acc;
}

// node counts per level: exclusive-scan id of the last entry + its own head flag
__global__ void pack_totals_kernel(const int* h_leaf, const int* id_leaf, const int* h_top, const int* id_top, const int* h_mid,
                                   const int* id_mid, long long n, int* totals) {
  totals[0] = id_leaf[n - 1] + h_leaf[n - 1];
  totals[1] = id_top[n - 1] + h_top[n - 1];
  totals[2] = h_mid ? id_mid[n - 1] + h_mid[n - 1] : 0;
}

__global__ void pack_empty_kernel(int* pos, int n, int* pos_top) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pos[i] = 0;
  if (i == 0 && pos_top) { pos_top[0] = 0; pos_top[1] = 0; }
}

template <typename K>
static int radix_sort_pairs(const K* kin, K* kout, const unsigned* vin, unsigned* vout, long long n, int end_bit) {
  size_t tmp_bytes = 0;
  TB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kin, kout, vin, vout, n, 0, end_bit, stream()));
  void* tmp = nullptr;
  TB_TRY(scratch_alloc(&tmp, tmp_bytes ? tmp_bytes : 16));
  cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, n, 0, end_bit, stream());
  scratch_free(tmp);
  TB_CUDA(e);
  count_launch(1);
  return TACO_B200_OK;
}

// host copy-out of one result array (HOST result space) or hand-over of the device array (DEVICE space)
static int publish_array(void** slot, void* dptr, size_t bytes) {
  if (result_space() == TACO_B200_SPACE_DEVICE) { *slot = dptr; return TACO_B200_OK; }
  void* h = malloc(bytes ? bytes : 16);
  if (!h) return fail(TACO_B200_ERR_ALLOC, "pack: cannot allocate host result array");
  if (bytes) TB_TRY(d2h_fresh(h, dptr, bytes));
  *slot = h;
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_pack(taco_tensor_t* A, taco_tensor_t* coo) {
  TB_TRY(ensure_init());
  if (!A || !coo) return fail(TACO_B200_ERR_ARG, "pack: NULL tensor");
  const int order = A->order;
  if (order != coo->order || (order != 2 && order != 3)) return fail(TACO_B200_ERR_UNSUPPORTED, "pack: order 2 or 3 tensors only");
  // levels are in STORAGE order on both sides (TensorBase::pack permutes the coordinates by the format's mode ordering
  // before it calls the helper, src/tensor.cpp:393-414): level l holds mode mode_ordering[l], e.g. {1,0} for CSC
  int ldim[3] = {0, 0, 0};
  bool seen[3] = {false, false, false};
  for (int l = 0; l < order; l++) {
    const int m = A->mode_ordering[l];
    if (m < 0 || m >= order || seen[m]) return fail(TACO_B200_ERR_ARG, "pack: bad mode ordering");
    seen[m] = true;
    if (coo->mode_ordering[l] != m) return fail(TACO_B200_ERR_ARG, "pack: coordinate buffer and result disagree on the mode ordering");
    if (A->dimensions[m] != coo->dimensions[m] || A->dimensions[m] <= 0) return fail(TACO_B200_ERR_ARG, "pack: bad dimension of mode %d", m);
    ldim[l] = A->dimensions[m];
  }
  enum { CSR, DCSR, CSF } kind;
  if (order == 2 && A->mode_types[0] == taco_mode_dense && A->mode_types[1] == taco_mode_sparse) kind = CSR;
  else if (order == 2 && A->mode_types[0] == taco_mode_sparse && A->mode_types[1] == taco_mode_sparse) kind = DCSR;
  else if (order == 3 && A->mode_types[0] == taco_mode_sparse && A->mode_types[1] == taco_mode_sparse && A->mode_types[2] == taco_mode_sparse) kind = CSF;
  else return fail(TACO_B200_ERR_UNSUPPORTED, "pack: target format must be {Dense,Compressed}, {Compressed,Compressed} or {Compressed x3}");
  DType dt, dtc;
  TB_TRY(dtype_of(A, &dt));
  TB_TRY(dtype_of(coo, &dtc));
  if (dt != dtc) return fail(TACO_B200_ERR_FORMAT, "pack: mixed component types");
  if (!coo->indices || !coo->indices[0] || !coo->indices[0][0]) return fail(TACO_B200_ERR_ARG, "pack: coordinate buffer has no pos array");
  int32_t first = 0, n32 = 0;
  TB_TRY(read_i32((const int32_t*)coo->indices[0][0], &first));
  TB_TRY(read_i32((const int32_t*)coo->indices[0][0] + 1, &n32));
  const long long n = (long long)n32 - first;
  if (first != 0 || n < 0 || n > INT32_MAX - 65536) return fail(TACO_B200_ERR_ARG, "pack: bad coordinate count");
  const size_t es = dsize(dt);
  const int rows = ldim[0];
  const int b1 = bits_for(ldim[1]), b0 = bits_for(ldim[0]);

  int *d_pos_top = nullptr, *d_crd_top = nullptr, *d_pos_mid = nullptr, *d_crd_mid = nullptr, *d_pos_leaf = nullptr, *d_crd_leaf = nullptr;
  void* d_vals = nullptr;
  int nleaf = 0, nmid = 0, ntop = 0;
  if (kind != CSR) TB_TRY(device_result_alloc((void**)&d_pos_top, sizeof(int) * 2));

  if (n == 0) {
    const int npos = kind == CSR ? rows + 1 : 1;
    TB_TRY(device_result_alloc((void**)&d_pos_leaf, sizeof(int) * (size_t)npos));
    if (kind == CSF) TB_TRY(device_result_alloc((void**)&d_pos_mid, sizeof(int)));
    pack_empty_kernel<<<(npos + 255) / 256, 256, 0, stream()>>>(d_pos_leaf, npos, d_pos_top);
    if (kind == CSF) pack_empty_kernel<<<1, 32, 0, stream()>>>(d_pos_mid, 1, nullptr);
    count_launch(1);
    TB_TRY(device_result_alloc((void**)&d_crd_leaf, 16));
    TB_TRY(device_result_alloc(&d_vals, 16));
    if (kind != CSR) TB_TRY(device_result_alloc((void**)&d_crd_top, 16));
    if (kind == CSF) TB_TRY(device_result_alloc((void**)&d_crd_mid, 16));
  } else {
    In c0, c1, c2, vin;
    for (int l = 0; l < order; l++)
      if (!coo->indices[l] || !coo->indices[l][1]) return fail(TACO_B200_ERR_ARG, "pack: level %d has no coordinate array", l);
    if (!coo->vals) return fail(TACO_B200_ERR_ARG, "pack: coordinate buffer has no values");
    TB_TRY(c0.acquire(coo->indices[0][1], sizeof(int32_t) * (size_t)n));
    TB_TRY(c1.acquire(coo->indices[1][1], sizeof(int32_t) * (size_t)n));
    if (order == 3) TB_TRY(c2.acquire(coo->indices[2][1], sizeof(int32_t) * (size_t)n));
    TB_TRY(vin.acquire(coo->vals, es * (size_t)n));
    ProfScope ps("pack_coo");
    const unsigned grid = (unsigned)((n + 255) / 256);
    void *key_a = nullptr, *key_b = nullptr, *idx_a = nullptr, *idx_b = nullptr, *flags = nullptr;
    TB_TRY(scratch_alloc(&key_a, 8 * (size_t)n));
    TB_TRY(scratch_alloc(&key_b, 8 * (size_t)n));
    TB_TRY(scratch_alloc(&idx_a, 4 * (size_t)n));
    TB_TRY(scratch_alloc(&idx_b, 4 * (size_t)n));
    const unsigned* perm_in = nullptr;
    if (order == 3) {                          // least significant mode first (stable)
      pack_key32_kernel<<<grid, 256, 0, stream()>>>(c2.as<int>(), n, (unsigned*)key_a, (unsigned*)idx_a);
      TB_TRY(radix_sort_pairs<unsigned>((const unsigned*)key_a, (unsigned*)key_b, (const unsigned*)idx_a, (unsigned*)idx_b, n,
                                        bits_for(ldim[2])));
      perm_in = (const unsigned*)idx_b;
    }
    pack_keys_kernel<<<grid, 256, 0, stream()>>>(c0.as<int>(), c1.as<int>(), perm_in, b1, n, (unsigned long long*)key_a, (unsigned*)idx_a);
    // idx_b may be the input permutation of pack_keys_kernel (already consumed): sort (key_a, idx_a) -> (key_b, idx_b)
    TB_TRY(radix_sort_pairs<unsigned long long>((const unsigned long long*)key_a, (unsigned long long*)key_b, (const unsigned*)idx_a,
                                                (unsigned*)idx_b, n, b0 + b1));
    const unsigned long long* key = (const unsigned long long*)key_b;
    const unsigned* perm = (const unsigned*)idx_b;
    // heads + ids (re-using key_a as three int arrays of n entries ... it holds 8n bytes = two of them; flags holds the rest)
    TB_TRY(scratch_alloc(&flags, 4 * (size_t)n * 4));
    int* h_leaf = (int*)key_a;
    int* h_top = (int*)key_a + n;
    int* h_mid = order == 3 ? (int*)flags : nullptr;
    int* id_leaf = (int*)flags + n;
    int* id_top = (int*)flags + 2 * n;
    int* id_mid = order == 3 ? (int*)flags + 3 * n : nullptr;
    pack_heads_kernel<<<grid, 256, 0, stream()>>>(key, order == 3 ? c2.as<int>() : nullptr, perm, b1, n, h_leaf, h_mid, h_top);
    count_launch(2 + (order == 3));
    TB_TRY(exclusive_scan_i32(h_leaf, id_leaf, n));
    TB_TRY(exclusive_scan_i32(h_top, id_top, n));
    if (order == 3) TB_TRY(exclusive_scan_i32(h_mid, id_mid, n));
    int last[3] = {0, 0, 0};
    {
      void* totals = nullptr;
      TB_TRY(scratch_alloc(&totals, sizeof(int) * 4));
      pack_totals_kernel<<<1, 1, 0, stream()>>>(h_leaf, id_leaf, h_top, id_top, h_mid, id_mid, n, (int*)totals);
      const int rc = read_back(last, totals, sizeof(int) * 3);
      scratch_free(totals);
      TB_TRY(rc);
    }
    nleaf = last[0]; ntop = last[1]; nmid = last[2];
    // result arrays
    const int npos_leaf = kind == CSR ? rows + 1 : (kind == DCSR ? ntop + 1 : nmid + 1);
    TB_TRY(device_result_alloc((void**)&d_pos_leaf, sizeof(int) * (size_t)npos_leaf));
    TB_TRY(device_result_alloc((void**)&d_crd_leaf, sizeof(int) * (size_t)nleaf));
    TB_TRY(device_result_alloc(&d_vals, es * (size_t)nleaf));
    if (kind != CSR) TB_TRY(device_result_alloc((void**)&d_crd_top, sizeof(int) * (size_t)ntop));
    if (kind == CSF) {
      TB_TRY(device_result_alloc((void**)&d_pos_mid, sizeof(int) * ((size_t)ntop + 1)));
      TB_TRY(device_result_alloc((void**)&d_crd_mid, sizeof(int) * (size_t)nmid));
    }
    void* run_start = nullptr;
    TB_TRY(scratch_alloc(&run_start, sizeof(int) * ((size_t)nleaf + 1)));
    PackOut o{d_pos_top, d_crd_top, d_pos_mid, d_crd_mid, d_pos_leaf, d_crd_leaf, (int*)run_start};
    pack_fill_kernel<<<grid, 256, 0, stream()>>>(key, order == 3 ? c2.as<int>() : nullptr, perm, b1, n, h_leaf, h_mid, h_top, id_leaf,
                                                id_mid, id_top, rows, nleaf, nmid, ntop, o);
    const unsigned vgrid = (unsigned)((nleaf + 255) / 256);
    if (dt == DType::F64) pack_vals_kernel<double><<<vgrid, 256, 0, stream()>>>(vin.as<double>(), perm, (const int*)run_start, nleaf, (double*)d_vals);
    else pack_vals_kernel<float><<<vgrid, 256, 0, stream()>>>(vin.as<float>(), perm, (const int*)run_start, nleaf, (float*)d_vals);
    count_launch(2);
    TB_CUDA(cudaGetLastError());
    scratch_free(run_start); scratch_free(flags); scratch_free(idx_b); scratch_free(idx_a); scratch_free(key_b); scratch_free(key_a);
  }

  // hand the arrays over in the configured result space
  const int last_level = order - 1;
  if (!A->indices) return fail(TACO_B200_ERR_ARG, "pack: result tensor has no indices table");
  const size_t npos_leaf = kind == CSR ? (size_t)rows + 1 : (kind == DCSR ? (size_t)ntop + 1 : (size_t)nmid + 1);
  void* p = nullptr;
  TB_TRY(publish_array(&p, d_pos_leaf, sizeof(int) * npos_leaf)); A->indices[last_level][0] = (uint8_t*)p;
  TB_TRY(publish_array(&p, d_crd_leaf, sizeof(int) * (size_t)nleaf)); A->indices[last_level][1] = (uint8_t*)p;
  TB_TRY(publish_array(&p, d_vals, es * (size_t)nleaf)); A->vals = (uint8_t*)p;
  if (kind != CSR) {
    TB_TRY(publish_array(&p, d_pos_top, sizeof(int) * 2)); A->indices[0][0] = (uint8_t*)p;
    TB_TRY(publish_array(&p, d_crd_top, sizeof(int) * (size_t)ntop)); A->indices[0][1] = (uint8_t*)p;
  }
  if (kind == CSF) {
    TB_TRY(publish_array(&p, d_pos_mid, sizeof(int) * ((size_t)ntop + 1))); A->indices[1][0] = (uint8_t*)p;
    TB_TRY(publish_array(&p, d_crd_mid, sizeof(int) * (size_t)nmid)); A->indices[1][1] = (uint8_t*)p;
  }
  A->vals_size = nleaf;
  if (result_space() != TACO_B200_SPACE_DEVICE) {
    TB_CUDA(cudaStreamSynchronize(stream()));
    device_result_free(d_pos_leaf); device_result_free(d_crd_leaf); device_result_free(d_vals);
    device_result_free(d_pos_top); device_result_free(d_crd_top); device_result_free(d_pos_mid); device_result_free(d_crd_mid);
  }
  return finish_call();
}

int _shim_taco_b200_pack(void** p) { return taco_b200_pack((taco_tensor_t*)p[0], (taco_tensor_t*)p[1]); }

}  // extern "C"
