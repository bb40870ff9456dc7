// ingest.cu -- file -> packed tensor, parsed on the device.
//
// Replaces the reference's host readers for the two text formats its tests and command-line tool use
//   Matrix Market coordinate (.mtx / .ttx: `%%MatrixMarket matrix|tensor coordinate real general|symmetric`)
//       /root/reference/src/storage/file_io_mtx.cpp:39-150  (header, '%' comments, size line, nnz 1-based entries;
//       symmetric files mirror every off-diagonal entry)
//   FROSTT (.tns: one entry per line, 1-based coordinates then the value; order and dimensions inferred from the data)
//       /root/reference/src/storage/file_io_tns.cpp:39-96
// which parse with getline + strtol/strtod, `insert()` every entry into a coordinate buffer and then `pack()` (host qsort +
// JIT-compiled helper, src/tensor.cpp:295-463).  Here the file's bytes are read once into pinned memory and uploaded; the
// device splits them into lines (newline count per 256-byte chunk -> prefix sum -> line offsets), one thread parses one
// entry, a second prefix sum compacts the entries (and appends the mirrored ones of a symmetric matrix), and
// taco_b200_pack's radix sort builds the level arrays -- the COO never exists on the host.
// Values must equal strtod()'s bit for bit.  The device parser is exact for every numeral with at most 19 significant
// digits and a decimal exponent in [-19, 19]: significands below 2^53 with exponents in [-22, 22] take one correctly
// rounded multiply or divide by an exact power of ten (Clinger's fast path), the rest an exact 128-bit integer product or
// quotient rounded to nearest-even by hand.  Anything else (20+ digits, larger exponents, inf / nan) is re-parsed by
// strtod on the host copy of the file and patched in -- a handful of numerals, never the bulk.
#include <cctype>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace tb {

constexpr int ING_CHUNK = 256;      // bytes per thread in the line-split passes
constexpr int ING_MAX_ORDER = 3;

__constant__ double c_pow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

__constant__ unsigned long long c_pow10u[20] = {1ULL, 10ULL, 100ULL, 1000ULL, 10000ULL, 100000ULL, 1000000ULL, 10000000ULL, 100000000ULL,
    1000000000ULL, 10000000000ULL, 100000000000ULL, 1000000000000ULL, 10000000000000ULL, 100000000000000ULL, 1000000000000000ULL,
    10000000000000000ULL, 100000000000000000ULL, 1000000000000000000ULL, 10000000000000000000ULL};

// v * 2^e2 (+ a sticky remainder below v's last bit) rounded to the nearest double, ties to even
__device__ __forceinline__ double ing_round128(unsigned __int128 v, int e2, bool sticky) {
  const unsigned long long hi = (unsigned long long)(v >> 64), lo = (unsigned long long)v;
  const int msb = hi ? 127 - __clzll((long long)hi) : 63 - __clzll((long long)lo);
  if (msb <= 52) return ldexp((double)lo, e2);
  const int shift = msb - 52;
  unsigned long long m = (unsigned long long)(v >> shift);
  const unsigned __int128 rem = v & ((((unsigned __int128)1) << shift) - 1), half = ((unsigned __int128)1) << (shift - 1);
  if (rem > half || (rem == half && (sticky || (m & 1)))) m++;
  return ldexp((double)m, shift + e2);
}

// mant * 10^exp10 as the correctly rounded double; false when the numeral is outside the exact device range
__device__ __forceinline__ bool ing_to_double(unsigned long long mant, int exp10, double* out) {
  if (mant == 0) { *out = 0.0; return true; }
  if (mant <= (1ULL << 53) && exp10 >= -22 && exp10 <= 22) {
    *out = exp10 >= 0 ? (double)mant * c_pow10[exp10] : (double)mant / c_pow10[-exp10];
    return true;
  }
  if (exp10 >= 0 && exp10 <= 19) {
    *out = ing_round128((unsigned __int128)mant * c_pow10u[exp10], 0, false);
    return true;
  }
  if (exp10 < 0 && exp10 >= -19) {
    const int lz = __clzll((long long)mant);
    const unsigned __int128 num = ((unsigned __int128)(mant << lz)) << 64;      // mant * 2^(lz + 64)
    const unsigned __int128 d = c_pow10u[-exp10];
    *out = ing_round128(num / d, -(lz + 64), (num % d) != 0);
    return true;
  }
  return false;
}

__global__ void __launch_bounds__(256) ing_count_newlines_kernel(const char* __restrict__ buf, long long n, int* __restrict__ cnt) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lo = c * ING_CHUNK;
  if (lo >= n) return;
  const long long hi = lo + ING_CHUNK < n ? lo + ING_CHUNK : n;
  int k = 0;
  for (long long q = lo; q < hi; q++) k += (buf[q] == '\n');
  cnt[c] = k;
}

// line_start[1 + (index of the newline)] = byte after it; line 0 starts at byte 0
__global__ void __launch_bounds__(256)
ing_line_starts_kernel(const char* __restrict__ buf, long long n, const int* __restrict__ off, long long* __restrict__ line_start) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long lo = c * ING_CHUNK;
  if (c == 0) line_start[0] = 0;
  if (lo >= n) return;
  const long long hi = lo + ING_CHUNK < n ? lo + ING_CHUNK : n;
  int k = off[c];
  for (long long q = lo; q < hi; q++)
    if (buf[q] == '\n') line_start[1 + k++] = q + 1;
}

__device__ __forceinline__ bool ing_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// status per line: 0 = blank (skipped), 1 = parsed, 2 = parsed but the value needs strtod on the host, 3 = malformed
__global__ void __launch_bounds__(256)
ing_parse_kernel(const char* __restrict__ buf, long long n, const long long* __restrict__ line_start, long long nlines, int order,
                 int* __restrict__ c0, int* __restrict__ c1, int* __restrict__ c2, double* __restrict__ vals,
                 unsigned char* __restrict__ status, int* __restrict__ dim_max) {
  const long long ln = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= nlines) return;
  long long p = line_start[ln];
  const long long end = (ln + 1 < nlines ? line_start[ln + 1] : n + 1) - 1;      // the newline (or end of file)
  while (p < end && ing_space(buf[p])) p++;
  if (p >= end) { status[ln] = 0; return; }
  int coord[ING_MAX_ORDER] = {0, 0, 0};
  bool bad = false;
  for (int m = 0; m < order; m++) {
    while (p < end && ing_space(buf[p])) p++;
    bool neg = false;
    if (p < end && (buf[p] == '-' || buf[p] == '+')) { neg = buf[p] == '-'; p++; }
    long long v = 0;
    int digits = 0;
    while (p < end && buf[p] >= '0' && buf[p] <= '9') { if (v < (1LL << 40)) v = v * 10 + (buf[p] - '0'); p++; digits++; }
    if (digits == 0 || neg || v < 1 || v > INT_MAX) bad = true;                     // coordinates are 1-based
    coord[m] = (int)(v - 1);
  }
  while (p < end && ing_space(buf[p])) p++;
  // value: [sign] digits [. digits] [e|E [sign] digits]
  bool neg = false, slow = false;
  if (p < end && (buf[p] == '-' || buf[p] == '+')) { neg = buf[p] == '-'; p++; }
  unsigned long long mant = 0;
  int sig = 0, dropped = 0, frac = 0, digits = 0;
  bool dot = false;
  while (p < end) {
    const char c = buf[p];
    if (c >= '0' && c <= '9') {
      digits++;
      if (mant == 0 && c == '0') { if (dot) frac++; }
      else if (sig < 19) { mant = mant * 10 + (unsigned)(c - '0'); sig++; if (dot) frac++; }
      else { dropped++; if (c != '0') slow = true; if (dot) {} else frac--; }
      p++;
    } else if (c == '.' && !dot) { dot = true; p++; }
    else break;
  }
  int exp10 = 0;
  if (p < end && (buf[p] == 'e' || buf[p] == 'E' || buf[p] == 'd' || buf[p] == 'D')) {
    if (buf[p] == 'd' || buf[p] == 'D') slow = true;      // Fortran exponent: strtod stops there; let the host decide
    long long q = p + 1;
    bool eneg = false;
    if (q < end && (buf[q] == '-' || buf[q] == '+')) { eneg = buf[q] == '-'; q++; }
    int ed = 0, ev = 0;
    while (q < end && buf[q] >= '0' && buf[q] <= '9') { if (ev < 100000) ev = ev * 10 + (buf[q] - '0'); q++; ed++; }
    if (ed > 0) { exp10 = eneg ? -ev : ev; p = q; }
  }
  if (digits == 0) { slow = true; }                        // "inf", "nan", hex floats, or garbage: strtod's call
  (void)dropped;
  exp10 -= frac;
  double v = 0.0;
  if (!slow && !ing_to_double(mant, exp10, &v)) slow = true;
  if (neg) v = -v;
  if (bad) { status[ln] = 3; return; }
  c0[ln] = coord[0];
  if (order > 1) c1[ln] = coord[1];
  if (order > 2) c2[ln] = coord[2];
  vals[ln] = v;
  status[ln] = slow ? 2 : 1;
  for (int m = 0; m < order; m++) atomicMax(dim_max + m, coord[m] + 1);
}

// keep[ln] = the line is an entry to be kept (MTX: only the first `limit` entries count, as the reference inserts only nnz);
// mirror[ln] = it also yields a transposed entry (symmetric matrices: off-diagonal entries)
__global__ void __launch_bounds__(256)
ing_flags_kernel(const unsigned char* __restrict__ status, const int* __restrict__ entry_idx, long long nlines, long long limit,
                 bool symm, const int* __restrict__ c0, const int* __restrict__ c1, int* __restrict__ keep, int* __restrict__ mirror,
                 int* __restrict__ counters) {
  const long long ln = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= nlines) return;
  const unsigned char st = status[ln];
  if (st == 3) atomicAdd(counters + 0, 1);
  const bool k = (st == 1 || st == 2) && (entry_idx == nullptr || entry_idx[ln] < limit);
  if (keep) keep[ln] = k;
  if (mirror) mirror[ln] = k && symm && c0[ln] != c1[ln];
  if (k && st == 2) atomicAdd(counters + 1, 1);
}

__global__ void __launch_bounds__(256)
ing_entry_flag_kernel(const unsigned char* __restrict__ status, long long nlines, int* __restrict__ isentry) {
  const long long ln = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ln < nlines) isentry[ln] = status[ln] == 1 || status[ln] == 2;
}

template <typename T>
__global__ void ing_patch_kernel(const int* __restrict__ slots, const double* __restrict__ v, int n, T* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && slots[i] >= 0) out[slots[i]] = (T)v[i >> 1];
}

template <typename T>
__global__ void __launch_bounds__(256)
ing_compact_kernel(long long nlines, int order, const int* __restrict__ keep_off, const int* __restrict__ keep, const int* __restrict__ mir_off,
                   const int* __restrict__ mirror, int nkept, const int* __restrict__ c0, const int* __restrict__ c1,
                   const int* __restrict__ c2, const double* __restrict__ vals, const unsigned char* __restrict__ status,
                   int* __restrict__ o0, int* __restrict__ o1, int* __restrict__ o2, T* __restrict__ ov, long long* __restrict__ slow_lines,
                   int* __restrict__ slow_slots, int* __restrict__ counters) {
  const long long ln = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ln >= nlines || !keep[ln]) return;
  const int e = keep_off[ln];
  o0[e] = c0[ln];
  if (order > 1) o1[e] = c1[ln];
  if (order > 2) o2[e] = c2[ln];
  ov[e] = (T)vals[ln];
  int m = -1;
  if (mirror && mirror[ln]) {
    m = nkept + mir_off[ln];
    o0[m] = c1[ln];
    o1[m] = c0[ln];
    ov[m] = (T)vals[ln];
  }
  if (status[ln] == 2) {                  // value to be re-parsed by strtod on the host: remember the line and its slots
    const int s = atomicAdd(counters + 2, 1);
    slow_lines[s] = ln;
    slow_slots[2 * s] = e;
    slow_slots[2 * s + 1] = m;
  }
}

static bool ends_with(const std::string& s, const char* suffix) {
  const size_t n = strlen(suffix);
  if (s.size() < n) return false;
  for (size_t i = 0; i < n; i++)
    if (tolower((unsigned char)s[s.size() - n + i]) != suffix[i]) return false;
  return true;
}

struct HostFile {
  char* data = nullptr;
  size_t size = 0;
  ~HostFile() { if (data) cudaFreeHost(data); }
};

static int load_file(const char* path, HostFile* f) {
  FILE* fp = fopen(path, "rb");
  if (!fp) return fail(TACO_B200_ERR_ARG, "read: cannot open '%s': %s", path, strerror(errno));
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (sz < 0) { fclose(fp); return fail(TACO_B200_ERR_ARG, "read: cannot size '%s'", path); }
  f->size = (size_t)sz;
  if (cudaHostAlloc((void**)&f->data, f->size + 16, cudaHostAllocDefault) != cudaSuccess) {
    fclose(fp);
    cudaGetLastError();
    return fail(TACO_B200_ERR_ALLOC, "read: cannot allocate %zu bytes of pinned memory", f->size);
  }
  const size_t got = fread(f->data, 1, f->size, fp);
  fclose(fp);
  if (got != f->size) return fail(TACO_B200_ERR_ARG, "read: short read of '%s'", path);
  f->data[f->size] = 0;
  return TACO_B200_OK;
}

// Matrix Market header (file_io_mtx.cpp:39-73, 86-112): banner, '%' comment lines, then the size line  d1 d2 [..] nnz.
static int parse_mtx_header(const HostFile& f, std::vector<int>* dims, long long* nnz, bool* symm, size_t* body) {
  size_t p = 0;
  auto next_line = [&](std::string* line) {
    if (p >= f.size) return false;
    size_t e = p;
    while (e < f.size && f.data[e] != '\n') e++;
    *line = std::string(f.data + p, e - p);
    p = e < f.size ? e + 1 : e;
    return true;
  };
  std::string line;
  if (!next_line(&line)) return fail(TACO_B200_ERR_ARG, "read: empty Matrix Market file");
  char head[64] = "", type[64] = "", fmt[64] = "", field[64] = "", symmetry[64] = "";
  sscanf(line.c_str(), "%63s %63s %63s %63s %63s", head, type, fmt, field, symmetry);
  if (strcmp(head, "%%MatrixMarket") != 0) return fail(TACO_B200_ERR_ARG, "read: unknown header of MatrixMarket");
  if (strcmp(type, "matrix") != 0 && strcmp(type, "tensor") != 0) return fail(TACO_B200_ERR_ARG, "read: unknown type of MatrixMarket");
  if (strcmp(field, "real") != 0) return fail(TACO_B200_ERR_UNSUPPORTED, "read: MatrixMarket field '%s' not available (real only, as in taco)", field);
  if (strcmp(symmetry, "general") != 0 && strcmp(symmetry, "symmetric") != 0)
    return fail(TACO_B200_ERR_UNSUPPORTED, "read: MatrixMarket symmetry '%s' not available", symmetry);
  if (strcmp(fmt, "coordinate") != 0)
    return fail(TACO_B200_ERR_UNSUPPORTED, "read: MatrixMarket format '%s' is not on the GPU path (coordinate files only)", fmt);
  *symm = strcmp(symmetry, "symmetric") == 0;
  bool have = false;
  while (next_line(&line)) {
    size_t q = 0;
    while (q < line.size() && isspace((unsigned char)line[q])) q++;
    if (q < line.size() && line[q] == '%') continue;
    have = true;
    break;
  }
  if (!have) return fail(TACO_B200_ERR_ARG, "read: Matrix Market file has no size line");
  std::vector<long long> nums;
  char* lp = (char*)line.c_str();
  while (true) {
    char* e = nullptr;
    const unsigned long long v = strtoull(lp, &e, 10);
    if (e == lp || v == 0) break;                    // the reference's loop also stops at the first 0 / non-number
    nums.push_back((long long)v);
    lp = e;
  }
  if (nums.size() < 2) return fail(TACO_B200_ERR_ARG, "read: malformed Matrix Market size line");
  *nnz = nums.back();
  nums.pop_back();
  for (long long d : nums) {
    if (d > INT_MAX) return fail(TACO_B200_ERR_ARG, "read: dimension exceeds INT_MAX");
    dims->push_back((int)d);
  }
  if (*symm && dims->size() != 2) return fail(TACO_B200_ERR_ARG, "read: symmetry only available for matrices");
  *body = p;
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

__global__ void ing_gather_starts_kernel(const long long* __restrict__ line_start, const long long* __restrict__ lines, int n,
                                         long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = line_start[lines[i]];
}

extern "C" int taco_b200_read(const char* path, taco_tensor_t* A) {
  TB_TRY(ensure_init());
  if (!path || !A) return fail(TACO_B200_ERR_ARG, "read: NULL argument");
  const std::string name(path);
  const bool mtx = ends_with(name, ".mtx") || ends_with(name, ".ttx");
  const bool tns = ends_with(name, ".tns");
  if (!mtx && !tns) return fail(TACO_B200_ERR_UNSUPPORTED, "read: '%s': only .mtx / .ttx (Matrix Market coordinate) and .tns (FROSTT) are read on the device", path);
  DType dt;
  TB_TRY(dtype_of(A, &dt));
  const int order = A->order;
  if (order < 2 || order > ING_MAX_ORDER) return fail(TACO_B200_ERR_FORMAT, "read: order-%d tensors are not packed on the device (2 or 3)", order);
  HostFile f;
  TB_TRY(load_file(path, &f));
  std::vector<int> dims;
  long long limit = LLONG_MAX;
  bool symm = false;
  size_t body = 0;
  if (mtx) {
    TB_TRY(parse_mtx_header(f, &dims, &limit, &symm, &body));
    if ((int)dims.size() != order) return fail(TACO_B200_ERR_ARG, "read: the file holds an order-%zu tensor, the result tensor has order %d", dims.size(), order);
  } else {
    // the order is the token count of the first line minus one (file_io_tns.cpp:51-53)
    size_t e = 0;
    int tokens = 0;
    bool in = false;
    while (e < f.size && f.data[e] != '\n') { const bool sp = isspace((unsigned char)f.data[e]); if (!sp && !in) tokens++; in = !sp; e++; }
    if (tokens - 1 != order) return fail(TACO_B200_ERR_ARG, "read: the file holds an order-%d tensor, the result tensor has order %d", tokens - 1, order);
  }
  const long long nbytes = (long long)(f.size - body);
  if (nbytes >= INT_MAX)      // line numbers and entry counts are int32 (as the tensor's positions are)
    return fail(TACO_B200_ERR_UNSUPPORTED, "read: '%s': files of 2 GiB and more are not read on the device", path);
  // ---- device: bytes -> lines -> entries ------------------------------------------------------------------------
  cudaStream_t st = stream();
  PipelineGuard g;                                   // releases every scratch buffer on all exit paths
  void *dbuf = nullptr, *dcnt = nullptr, *dls = nullptr;
  const long long nchunks = (nbytes + ING_CHUNK - 1) / ING_CHUNK;
  TB_TRY(g.alloc(&dbuf, (size_t)nbytes + 16));
  TB_TRY(g.alloc(&dcnt, sizeof(int) * (size_t)(nchunks + 1)));
  if (nbytes) TB_CUDA(cudaMemcpyAsync(dbuf, f.data + body, (size_t)nbytes, cudaMemcpyHostToDevice, st));
  TB_CUDA(cudaMemsetAsync(dcnt, 0, sizeof(int) * (size_t)(nchunks + 1), st));
  const unsigned cgrid = (unsigned)((nchunks + 255) / 256);
  if (nchunks) ing_count_newlines_kernel<<<cgrid, 256, 0, st>>>((const char*)dbuf, nbytes, (int*)dcnt);
  TB_TRY(exclusive_scan_i32((const int*)dcnt, (int*)dcnt, nchunks + 1));
  int newlines = 0;
  TB_TRY(read_back(&newlines, (int*)dcnt + nchunks, sizeof(int)));
  const long long nlines = (long long)newlines + 1;  // the text after the last newline is a line too (possibly blank)
  TB_TRY(g.alloc(&dls, sizeof(long long) * (size_t)nlines));
  if (nchunks) ing_line_starts_kernel<<<cgrid, 256, 0, st>>>((const char*)dbuf, nbytes, (const int*)dcnt, (long long*)dls);
  else TB_CUDA(cudaMemsetAsync(dls, 0, sizeof(long long), st));
  void *c0 = nullptr, *c1 = nullptr, *c2 = nullptr, *pv = nullptr, *stt = nullptr, *counters = nullptr, *flag_a = nullptr, *flag_b = nullptr, *eidx = nullptr;
  TB_TRY(g.alloc(&c0, sizeof(int) * (size_t)nlines));
  TB_TRY(g.alloc(&c1, sizeof(int) * (size_t)nlines));
  if (order > 2) TB_TRY(g.alloc(&c2, sizeof(int) * (size_t)nlines));
  TB_TRY(g.alloc(&pv, sizeof(double) * (size_t)nlines));
  TB_TRY(g.alloc(&stt, (size_t)nlines));
  TB_TRY(g.alloc(&counters, sizeof(int) * 8));       // [0] malformed lines, [1] slow values kept, [2] slow list cursor, [4..6] max coordinate + 1
  TB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * 8, st));
  const unsigned lgrid = (unsigned)((nlines + 255) / 256);
  ing_parse_kernel<<<lgrid, 256, 0, st>>>((const char*)dbuf, nbytes, (const long long*)dls, nlines, order, (int*)c0, (int*)c1, (int*)c2,
                                          (double*)pv, (unsigned char*)stt, (int*)counters + 4);
  TB_TRY(g.alloc(&flag_a, sizeof(int) * (size_t)(nlines + 1)));
  TB_TRY(g.alloc(&flag_b, sizeof(int) * (size_t)(nlines + 1)));
  if (mtx) {                                         // entry number of every line, so that only the first nnz entries are kept
    TB_TRY(g.alloc(&eidx, sizeof(int) * (size_t)(nlines + 1)));
    ing_entry_flag_kernel<<<lgrid, 256, 0, st>>>((const unsigned char*)stt, nlines, (int*)eidx);
    TB_TRY(exclusive_scan_i32((const int*)eidx, (int*)eidx, nlines + 1));
  }
  TB_CUDA(cudaMemsetAsync((int*)flag_a + nlines, 0, sizeof(int), st));
  TB_CUDA(cudaMemsetAsync((int*)flag_b + nlines, 0, sizeof(int), st));
  ing_flags_kernel<<<lgrid, 256, 0, st>>>((const unsigned char*)stt, (const int*)eidx, nlines, limit, symm, (const int*)c0, (const int*)c1,
                                          (int*)flag_a, (int*)flag_b, (int*)counters);
  void *koff = nullptr, *moff = nullptr;
  TB_TRY(g.alloc(&koff, sizeof(int) * (size_t)(nlines + 1)));
  TB_TRY(g.alloc(&moff, sizeof(int) * (size_t)(nlines + 1)));
  TB_TRY(exclusive_scan_i32((const int*)flag_a, (int*)koff, nlines + 1));
  TB_TRY(exclusive_scan_i32((const int*)flag_b, (int*)moff, nlines + 1));
  count_launch(6);
  int h[8], nk = 0, nm = 0;
  TB_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(&nk, (int*)koff + nlines, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaMemcpyAsync(&nm, (int*)moff + nlines, sizeof(int), cudaMemcpyDeviceToHost, st));
  TB_CUDA(cudaStreamSynchronize(st));
  if (h[0] > 0) return fail(TACO_B200_ERR_ARG, "read: '%s' has %d malformed entry lines (coordinates must be positive integers)", path, h[0]);
  if (mtx && (long long)nk < limit) return fail(TACO_B200_ERR_ARG, "read: '%s' announces %lld entries but holds %d", path, limit, nk);
  const long long n = (long long)nk + nm;
  if (n > INT_MAX - 65536) return fail(TACO_B200_ERR_ARG, "read: too many entries for int32 positions");
  // dimensions: the size line (mtx) or the largest coordinate per mode (tns); a caller-supplied dimension may be larger
  for (int m = 0; m < order; m++) {
    const int fromfile = mtx ? dims[m] : h[4 + m];
    if (mtx && h[4 + m] > dims[m]) return fail(TACO_B200_ERR_ARG, "read: a coordinate of mode %d exceeds the dimension in the size line", m);
    if (A->dimensions[m] <= 0) A->dimensions[m] = fromfile;
    else if (A->dimensions[m] < fromfile) return fail(TACO_B200_ERR_ARG, "read: dimension %d of the result tensor (%d) is smaller than the file's (%d)", m, A->dimensions[m], fromfile);
  }
  for (int l = 0; l < order; l++)
    if (A->mode_types[l] == taco_mode_dense && A->indices && A->indices[l] && A->indices[l][0])
      *(int32_t*)A->indices[l][0] = A->dimensions[A->mode_ordering[l]];
  // ---- compact into the COO arrays (file order; mirrored entries of a symmetric matrix after them) -----------------
  void *o[3] = {nullptr, nullptr, nullptr}, *ov = nullptr, *slow_lines = nullptr, *slow_slots = nullptr;
  const size_t es = dsize(dt), cap = (size_t)(n > 0 ? n : 1);
  for (int m = 0; m < order; m++) TB_TRY(g.alloc(&o[m], sizeof(int) * cap));
  TB_TRY(g.alloc(&ov, es * cap));
  const int nslow = h[1];
  TB_TRY(g.alloc(&slow_lines, sizeof(long long) * (size_t)(nslow + 1)));
  TB_TRY(g.alloc(&slow_slots, sizeof(int) * 2 * (size_t)(nslow + 1)));
  if (dt == DType::F64)
    ing_compact_kernel<double><<<lgrid, 256, 0, st>>>(nlines, order, (const int*)koff, (const int*)flag_a, (const int*)moff, symm ? (const int*)flag_b : nullptr,
        nk, (const int*)c0, (const int*)c1, (const int*)c2, (const double*)pv, (const unsigned char*)stt, (int*)o[0], (int*)o[1], (int*)o[2], (double*)ov,
        (long long*)slow_lines, (int*)slow_slots, (int*)counters);
  else
    ing_compact_kernel<float><<<lgrid, 256, 0, st>>>(nlines, order, (const int*)koff, (const int*)flag_a, (const int*)moff, symm ? (const int*)flag_b : nullptr,
        nk, (const int*)c0, (const int*)c1, (const int*)c2, (const double*)pv, (const unsigned char*)stt, (int*)o[0], (int*)o[1], (int*)o[2], (float*)ov,
        (long long*)slow_lines, (int*)slow_slots, (int*)counters);
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  if (nslow > 0) {                                   // numerals outside the exact fast path: strtod on the host copy of the file
    std::vector<long long> lines(nslow), starts(nslow);
    std::vector<int> slots(2 * (size_t)nslow);
    TB_TRY(read_back(lines.data(), slow_lines, sizeof(long long) * (size_t)nslow));
    TB_TRY(read_back(slots.data(), slow_slots, sizeof(int) * 2 * (size_t)nslow));
    void* dstarts = nullptr;                         // the byte offsets of those lines: one gather, one read-back
    TB_TRY(g.alloc(&dstarts, sizeof(long long) * (size_t)nslow));
    ing_gather_starts_kernel<<<(nslow + 255) / 256, 256, 0, st>>>((const long long*)dls, (const long long*)slow_lines, nslow, (long long*)dstarts);
    count_launch(1);
    TB_TRY(read_back(starts.data(), dstarts, sizeof(long long) * (size_t)nslow));
    std::vector<double> fixed(nslow);
    for (int s = 0; s < nslow; s++) {
      char* lp = f.data + body + starts[s];
      for (int m = 0; m < order; m++) strtol(lp, &lp, 10);
      fixed[s] = strtod(lp, &lp);
    }
    void* dfixed = nullptr;                          // one upload, one scatter: slot 2s / 2s+1 (entry, mirrored entry) <- value s
    TB_TRY(g.alloc(&dfixed, sizeof(double) * (size_t)nslow));
    TB_CUDA(cudaMemcpyAsync(dfixed, fixed.data(), sizeof(double) * (size_t)nslow, cudaMemcpyHostToDevice, st));
    if (dt == DType::F64) ing_patch_kernel<double><<<(2 * nslow + 255) / 256, 256, 0, st>>>((const int*)slow_slots, (const double*)dfixed, 2 * nslow, (double*)ov);
    else ing_patch_kernel<float><<<(2 * nslow + 255) / 256, 256, 0, st>>>((const int*)slow_slots, (const double*)dfixed, 2 * nslow, (float*)ov);
    count_launch(1);
    TB_CUDA(cudaStreamSynchronize(st));             // `fixed` is a local
  }
  // ---- the coordinate-buffer tensor taco_b200_pack takes: level l holds the coordinates of mode mode_ordering[l] ----------
  int32_t coo_pos[2] = {0, (int32_t)n};
  uint8_t* lvl[ING_MAX_ORDER][2];
  uint8_t** idx[ING_MAX_ORDER];
  taco_mode_t types[ING_MAX_ORDER];
  for (int l = 0; l < order; l++) {
    lvl[l][0] = l == 0 ? (uint8_t*)coo_pos : nullptr;
    lvl[l][1] = (uint8_t*)o[A->mode_ordering[l]];
    idx[l] = lvl[l];
    types[l] = taco_mode_sparse;
  }
  taco_tensor_t coo;
  memset(&coo, 0, sizeof(coo));
  coo.order = order;
  coo.dimensions = A->dimensions;
  coo.csize = A->csize;
  coo.mode_ordering = A->mode_ordering;
  coo.mode_types = types;
  coo.indices = idx;
  coo.vals = (uint8_t*)ov;
  coo.vals_size = (int32_t)n;
  return taco_b200_pack(A, &coo);
}

extern "C" int _shim_taco_b200_read(void** p) { return taco_b200_read((const char*)p[0], (taco_tensor_t*)p[1]); }
