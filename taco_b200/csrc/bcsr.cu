// bcsr.cu -- blocked SpMV / SpMM over the reference's blocked format {Dense, Compressed, Dense, Dense}
//   a(i,j)   = A(i,k,j,l) * c(k,l)        (the reference's `bspmv` statement, /root/reference/test/tests-expr_storage.cpp:939-960)
//   C(i,j,m) = A(i,k,j,l) * B(k,l,m)      (blocked SpMM: the one hot-path contraction that is DENSE inside a block)
// A is order 4 over (block row i, block column k, row in block j, column in block l): pos = int32[Mb+1] and crd over the
// block columns, vals = [stored blocks][br][bc]; B = (Nb, bc, K) and C = (Mb, br, K) are dense, i.e. row-major
// (Nb*bc) x K and (Mb*br) x K matrices.
//
// Kernels:
//   * bspmm_rows_kernel (fp32 / fp64, any block shape): CUDA cores, the reference's summation order (block by block,
//     then column l inside the block, separate multiply and add) -> bit-identical to the reference's generated C.
//   * bspmm_tma_kernel (DEFAULT for fp32, 16x16 and 32x32 blocks, K % 4 == 0) and bspmm_tc_kernel (its register-staged
//     predecessor, TACO_B200_BSPMM_VARIANT=4): tcgen05 tensor cores.  The block product is computed transposed,
//     D[k', i2] += Bt[k', j2] * At[j2, i2], so the MMA's M dimension is the 128-wide tile of dense columns (always
//     full) and its N dimension is the block height.  fp32 inputs are split into two TF32 terms (x = hi + lo, hi = top
//     19 bits) and the three products hi*hi, hi*lo, lo*hi accumulate in fp32 in tensor memory (two MMAs per K step: the
//     A block is ONE operand of 2*br rows [hi ; lo]), which keeps the result within the fp32 tolerance of the north
//     star (1e-5 relative; a single TF32 pass would be 1e-3).  Warp-specialised: TMA (or producer) warps -> mbarrier
//     rings -> 1 MMA warp (one thread issues tcgen05.mma, tcgen05.commit frees stages) -> 4 epilogue warps
//     (tcgen05.ld, coalesced 128-byte stores); two accumulators in TMEM so the epilogue of one block row overlaps the
//     MMAs of the next.  Measured at the bench config (32768^2 blocks of 32x32, 16 per block row, K = 128):
//     CUDA cores 21.3 ms -> register-staged tcgen05 2.64 ms -> TMA-fed 1.60 ms = 85.6 TFLOP/s with 10.2 GB of DRAM
//     traffic, i.e. 6.4 TB/s = the measured HBM peak (profiles/r01_bspmm.md).
// Algorithmic bytes per launch: 4(Mb+1) + nnzb*(4 + br*bc*es) + es*K*(Nb*bc + Mb*br).
#include <cuda.h>

#include <climits>
#include <type_traits>
#include <cstdlib>

#include "common.cuh"

namespace tb {

struct BcsrView { int32_t Mb, Nb, br, bc; int32_t* pos; int32_t* crd; void* vals; DType dt; };

static int view_bcsr(const taco_tensor_t* t, const char* name, BcsrView* v) {
  if (!t) return fail(TACO_B200_ERR_ARG, "%s: NULL tensor", name);
  if (t->order != 4 || t->mode_types[0] != taco_mode_dense || t->mode_types[1] != taco_mode_sparse ||
      t->mode_types[2] != taco_mode_dense || t->mode_types[3] != taco_mode_dense)
    return fail(TACO_B200_ERR_FORMAT, "%s: expected blocked CSR ({Dense,Compressed,Dense,Dense})", name);
  for (int l = 0; l < 4; l++)
    if (t->mode_ordering[l] != l) return fail(TACO_B200_ERR_FORMAT, "%s: blocked CSR needs mode ordering 0,1,2,3", name);
  v->Mb = t->dimensions[0]; v->Nb = t->dimensions[1]; v->br = t->dimensions[2]; v->bc = t->dimensions[3];
  if (v->Mb < 0 || v->Nb < 0 || v->br <= 0 || v->bc <= 0) return fail(TACO_B200_ERR_ARG, "%s: bad dimension", name);
  v->pos = t->indices && t->indices[1] ? (int32_t*)t->indices[1][0] : nullptr;
  v->crd = t->indices && t->indices[1] ? (int32_t*)t->indices[1][1] : nullptr;
  v->vals = t->vals;
  return dtype_of(t, &v->dt);
}

// number of stored blocks: pos[Mb] (device-resident tensors carry it in vals_size = blocks * br * bc)
static int bcsr_nnzb(const BcsrView& A, int32_t vals_size_hint, int32_t* nnzb) {
  if (!A.pos) return fail(TACO_B200_ERR_ARG, "blocked CSR operand has no pos array");
  if (vals_size_hint > 0 && trusts_vals_size(A.pos)) { *nnzb = vals_size_hint / (A.br * A.bc); return TACO_B200_OK; }
  return read_i32(A.pos + A.Mb, nnzb);
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-core kernels (reference order)
// ---------------------------------------------------------------------------------------------------------
// One thread per result row (i, j): tl = sum_l A[kA][j][l] * c[k][l] (scalar temporary), a[i,j] = a[i,j] + tl.
template <typename T>
__global__ void __launch_bounds__(256)
bspmv_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
             const T* __restrict__ c, T* __restrict__ a, int Mb, int br, int bc) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)Mb * br) return;
  const int i = (int)(t / br), j = (int)(t % br);
  T acc = T(0);
  for (int kA = __ldg(pos + i); kA < __ldg(pos + i + 1); kA++) {
    const T* blk = vals + ((size_t)kA * br + j) * bc;
    const T* cr = c + (size_t)__ldg(crd + kA) * bc;
    T tl = T(0);
    for (int l = 0; l < bc; l++) tl = tl + __ldg(blk + l) * __ldg(cr + l);
    acc = acc + tl;
  }
  a[t] = acc;
}

// Staged variant for br <= 32: a warp owns a block row.  Each stored block is read with coalesced 16-byte loads into a padded
// shared-memory tile (the whole block is one contiguous br*bc run in vals), c(k,:) likewise; lane j then forms tl for its
// row j in ascending l from shared memory and adds it -- the same operation order as bspmv_kernel and the reference, so the
// result is bit-identical.  The block stream is the only large array: algorithmic bytes = nnzb*(4 + br*bc*es) + small.
template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
bspmv_warp_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                  const T* __restrict__ c, T* __restrict__ a, int Mb, int br, int bc) {
  extern __shared__ __align__(16) unsigned char bspmv_smem[];
  constexpr int VEC = 16 / (int)sizeof(T);
  constexpr int EPL = 16;                                              // 16-byte pieces of a block per lane (<= 32 x 32 fp64)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = bc + 1;                                               // padded row: lane j reads column l of row j conflict-free
  T* tile = (T*)bspmv_smem + (size_t)warp * (br * ld + bc);
  T* sc = tile + br * ld;
  const int nel = br * bc;
  // (row, column) of this lane's piece t: incremental, no division in the loop (bc % VEC == 0, so a piece never straddles rows)
  const int step = 32 * VEC, dr = step / bc, dl = step - dr * bc;
  const int r0 = (lane * VEC) / bc, l0 = lane * VEC - r0 * bc;
  using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
  auto load_block = [&](int p, V (&x)[EPL], T& cv) {
    const T* blk = vals + (size_t)p * nel;
#pragma unroll
    for (int t = 0; t < EPL; t++) {
      const int e = (t * 32 + lane) * VEC;
      if (e < nel) {
        if constexpr (sizeof(T) == 4) x[t] = tbd::ldg_stream_f4(blk + e);
        else x[t] = tbd::ldg_stream_d2(blk + e);
      }
    }
    cv = T(0);
    const T* cr = c + (size_t)__ldg(crd + p) * bc;
    if (lane < bc) cv = __ldg(cr + lane);                               // bc <= 32 here; wider blocks take the second load below
  };
  for (int i = blockIdx.x * WARPS + warp; i < Mb; i += gridDim.x * WARPS) {
    T acc = T(0);
    const int p0 = __ldg(pos + i), p1 = __ldg(pos + i + 1);
    V cur[EPL], nxt[EPL];
    T ccur = T(0), cnxt = T(0);
    if (p0 < p1) load_block(p0, cur, ccur);
    for (int p = p0; p < p1; p++) {
      if (p + 1 < p1) load_block(p + 1, nxt, cnxt);                     // the next block is in flight while this one is reduced
      int r = r0, l = l0;
#pragma unroll
      for (int t = 0; t < EPL; t++) {
        if ((t * 32 + lane) * VEC < nel) {
          T* d = tile + r * ld + l;
          if constexpr (sizeof(T) == 4) { d[0] = cur[t].x; d[1] = cur[t].y; d[2] = cur[t].z; d[3] = cur[t].w; }
          else { d[0] = cur[t].x; d[1] = cur[t].y; }
        }
        l += dl; r += dr;
        if (l >= bc) { l -= bc; r++; }
      }
      if (lane < bc) sc[lane] = ccur;
      if (bc > 32) for (int q = 32 + lane; q < bc; q += 32) sc[q] = __ldg(c + (size_t)__ldg(crd + p) * bc + q);
      __syncwarp();
      if (lane < br) {
        const T* row = tile + lane * ld;
        T tl = T(0);
        for (int q = 0; q < bc; q++) tl = tl + row[q] * sc[q];
        acc = acc + tl;
      }
      __syncwarp();
#pragma unroll
      for (int t = 0; t < EPL; t++) cur[t] = nxt[t];
      ccur = cnxt;
    }
    if (lane < br) a[(size_t)i * br + lane] = acc;
  }
}

// CTA = one block row x one tile of TK dense columns.  Warp w owns the rows j = w, w + WARPS, ... (RJ accumulators per
// lane and column), lanes own columns.  The A block is staged in shared memory (broadcast reads), a row of B is read
// once per (block, l) and applied to all rows the warp owns.  Order per result: blocks ascending, then l ascending.
template <typename T, int WARPS, int RJ>
__global__ void __launch_bounds__(WARPS * 32)
bspmm_rows_kernel(const int* __restrict__ pos, const int* __restrict__ crd, const T* __restrict__ vals,
                  const T* __restrict__ B, T* __restrict__ C, int br, int bc, int K) {
  extern __shared__ __align__(16) unsigned char bs_smem[];
  T* sA = (T*)bs_smem;                                   // [br * bc]
  const int i = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m = blockIdx.y * 32 + lane;
  const bool active = m < K;
  const int p0 = __ldg(pos + i), p1 = __ldg(pos + i + 1);
  for (int j0 = 0; j0 < br; j0 += WARPS * RJ) {           // row groups (one pass when br <= WARPS * RJ)
    T acc[RJ];
#pragma unroll
    for (int r = 0; r < RJ; r++) acc[r] = T(0);
    for (int kA = p0; kA < p1; kA++) {
      __syncthreads();
      for (int q = threadIdx.x; q < br * bc; q += WARPS * 32) sA[q] = __ldg(vals + (size_t)kA * br * bc + q);
      __syncthreads();
      const T* bk = B + (size_t)__ldg(crd + kA) * bc * K + (active ? m : 0);
      for (int l = 0; l < bc; l++) {
        const T b = __ldg(bk + (size_t)l * K);
#pragma unroll
        for (int r = 0; r < RJ; r++) {
          const int j = j0 + w + r * WARPS;
          if (j < br) acc[r] = acc[r] + sA[j * bc + l] * b;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RJ; r++) {
      const int j = j0 + w + r * WARPS;
      if (j < br && active) C[((size_t)i * br + j) * K + m] = acc[r];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ---------------------------------------------------------------------------------------------------------
namespace tc {

constexpr int TILE_K = 128;                  // dense columns per CTA tile = the MMA's M
constexpr int EPI_WARPS = 4;
constexpr int MMA_WARP = EPI_WARPS;          // warp 4

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\n"
      "WAIT_DONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {   // arrives on `bar` when all MMAs issued so far have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], TF32 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor: start >> 4 at [0,14), leading byte offset
// >> 4 at [16,30), stride byte offset >> 4 at [32,46), version 1 at [46,48), layout type 0 at [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);
}

template <int BR, int BC>
struct Cfg {
  static_assert(BR % 16 == 0 && BR >= 16 && BR <= 64 && (BR & (BR - 1)) == 0, "block height = MMA N: multiple of 16");
  static_assert(BC >= 8 && BC <= 64 && (BC & (BC - 1)) == 0, "block width: power of two, multiple of the TF32 MMA K (8)");
  static constexpr int BT_BYTES = BC * TILE_K * 4;        // one B tile (BC rows x 128 columns), hi or lo
  static constexpr int AB_BYTES = BR * BC * 4;            // one A block, hi or lo
  static constexpr int STAGE_BYTES = 2 * BT_BYTES + 2 * AB_BYTES;
  // Both operands are K-major, no swizzle (measured with tools/tc_probe.cu: the MN-major / "transpose" bit yields zeros
  // for kind::tf32): core matrix = 8 rows x 16 bytes (4 j2); SBO = stride between 8-row groups, LBO = stride between
  // the two 16-byte chunks of one K = 8 step.  Chunk jc of row group g lives at (jc * groups + g) * 128.
  static constexpr int BT_SBO = 128;
  static constexpr int BT_LBO = (TILE_K / 8) * 128;       // B tile transposed: rows = dense columns k' (16 groups)
  static constexpr int AB_SBO = 128;
  static constexpr int AB_LBO = (2 * BR / 8) * 128;       // A block: ONE operand of 2 BR rows = [hi rows ; lo rows]
  static constexpr int ACC_COLS = 2 * BR;                 // accumulator: columns [0,BR) = hi*hi + lo*hi, [BR,2BR) = hi*lo
  static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B TF32, both K-major, N = BR, M = 128
  static constexpr uint32_t idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_K >> 4) << 24);
  }
  static constexpr uint32_t IDESC_N1 = idesc(BR), IDESC_N2 = idesc(2 * BR);
};

__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); lo.x = x.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); lo.y = x.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); lo.z = x.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); lo.w = x.w - hi.w;
}
__device__ __forceinline__ void sts_f4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// tcgen05 kernel, TMA-fed ("bspmm_tma_kernel")
// ---------------------------------------------------------------------------------------------------------
// Same math and the same accumulator / epilogue as bspmm_tc_kernel, but no operand passes through registers on its way
// in.  Measured with tools/tc_probe.cu on B200:
//   * kind::tf32 reads fp32 containers and IGNORES the low 13 mantissa bits (truncation), so the raw fp32 B tile is
//     already the "hi" operand; only lo = x - trunc(x) has to be computed;
//   * an MN-major tf32 operand must use the 128-byte swizzle with 32-byte atoms (descriptor layout type 1, Swizzle<2,5,2>
//     on byte addresses, LBO = stride between 32-column panels, SBO = 512 = four 128-byte rows) -- exactly what TMA
//     writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B and a {32 columns, BC rows} box.  So the B tile is NOT transposed:
//     rows j2 of B are the K rows of the MMA's A operand.
// Roles: 4 epilogue warps, 1 MMA warp, 1 TMA warp (per stored block: four tensor boxes of B + one 1-D bulk copy of the A
// block into the raw ring), CW converter warps (raw ring -> lo ring: Bt_lo in place-identical swizzled positions, the A
// block as the merged K-major [hi ; lo] operand).  Rings: RS raw stages (TMA in flight), LS lo stages.
namespace tma {

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_box_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// MN-major operand, 128-byte swizzle with 32-byte atoms (layout type 1)
__device__ __forceinline__ uint64_t smem_desc_mn32(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return smem_desc(addr, lbo, sbo) | (1ull << 61);
}

template <int BR, int BC>
struct TCfg {
  using CF = Cfg<BR, BC>;
  static constexpr int PANEL_BYTES = BC * 128;                      // one TMA box: BC rows x 32 dense columns
  static constexpr int BT_BYTES = 4 * PANEL_BYTES;                  // = CF::BT_BYTES
  static constexpr int AB_BYTES = BR * BC * 4;
  static constexpr int RAW_BYTES = BT_BYTES + AB_BYTES;             // raw stage: B tile (swizzled panels) + A block (row-major)
  static constexpr int LO_BYTES = BT_BYTES + 2 * AB_BYTES;          // lo stage: Bt_lo + merged [A_hi ; A_lo] (K-major, no swizzle)
  static_assert(RAW_BYTES % 1024 == 0 && LO_BYTES % 1024 == 0, "stages keep the 1024-byte swizzle phase");
  static constexpr uint32_t idesc(int n) { return CF::idesc(n) | (1u << 15); }     // A operand (the B tile) is MN-major
  // The tensor core adds into the fp32 accumulator with truncation, a bias of ~0.25 ulp of the running sum per MMA
  // (measured: 70 blocks of 32 columns, same-sign data -> 1.5e-5 relative).  A block row is therefore accumulated in
  // segments of at most 128 MMA steps; the epilogue adds the segments in fp32 (round to nearest), which keeps the bias
  // below 4e-6 of the row sum for any row length.
  static constexpr int SEG = 512 / BC;
};

template <int BR, int BC, int RS, int LS, int CW, int MINB>
__global__ void __launch_bounds__((EPI_WARPS + 2 + CW) * 32, MINB)
bspmm_tma_kernel(const __grid_constant__ CUtensorMap tmB, const int* __restrict__ pos, const int* __restrict__ crd,
                 const float* __restrict__ vals, float* __restrict__ C, int Mb, int K) {
  using CF = Cfg<BR, BC>;
  using TC = TCfg<BR, BC>;
  constexpr int TMA_WARP = EPI_WARPS + 1, CONV_WARP0 = EPI_WARPS + 2;
  extern __shared__ __align__(1024) unsigned char tc_smem[];
  const uint32_t smem0 = (smem_u32(tc_smem) + 1023u) & ~1023u;
  const uint32_t raw0 = smem0, lo0 = smem0 + RS * TC::RAW_BYTES;
  const uint32_t bars = lo0 + LS * TC::LO_BYTES;        // raw_full[RS] raw_empty[RS] lo_full[LS] lo_empty[LS] tfull[2] tempty[2] slot
  auto raw_full = [&](int s) { return bars + 8u * s; };
  auto raw_empty = [&](int s) { return bars + 8u * (RS + s); };
  auto lo_full = [&](int s) { return bars + 8u * (2 * RS + s); };
  auto lo_empty = [&](int s) { return bars + 8u * (2 * RS + LS + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * RS + 2 * LS + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * RS + 2 * LS + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * RS + 2 * LS + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.y * TILE_K;

  if (threadIdx.x == 0) {
    for (int s = 0; s < RS; s++) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), 1); }
    for (int s = 0; s < LS; s++) { mbar_init(lo_full(s), CW * 32); mbar_init(lo_empty(s), 1); }
    for (int a = 0; a < 2; a++) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS * 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(CF::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp == TMA_WARP && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp < EPI_WARPS) {
    // ===== epilogue (as bspmm_tc_kernel): TMEM lane = dense column k', TMEM column = row in block =====================
    const int col = k0 + warp * 32 + lane;
    uint32_t acc_it = 0;
    for (int i1 = blockIdx.x; i1 < Mb; i1 += gridDim.x) {
      const int n = __ldg(pos + i1 + 1) - __ldg(pos + i1);
      float* crow = C + (size_t)i1 * BR * K + col;
      if (n == 0) {
        if (col < K) for (int j = 0; j < BR; j++) crow[(size_t)j * K] = 0.0f;
        continue;
      }
      for (int b0 = 0; b0 < n; b0 += TC::SEG) {
      const uint32_t a = acc_it & 1, aph = (acc_it >> 1) & 1;
      mbar_wait(tfull_bar(a), aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + a * CF::ACC_COLS;
#pragma unroll
      for (int c = 0; c < BR; c += 16) {
        uint32_t v[16], x[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr + c) : "memory");
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
              "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15])
            : "r"(taddr + BR + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (c + 16 >= BR) {
          tc_fence_before();
          mbar_arrive(tempty_bar(a));
        }
        if (col < K) {
          float* p = crow + (size_t)c * K;
          if (b0 == 0) {
#pragma unroll
            for (int j = 0; j < 16; j++, p += K) *p = __uint_as_float(v[j]) + __uint_as_float(x[j]);
          } else {         // later segments of a long block row: fp32 add (round to nearest) onto what this thread stored before
#pragma unroll
            for (int j = 0; j < 16; j++, p += K) *p = *p + (__uint_as_float(v[j]) + __uint_as_float(x[j]));
          }
        }
      }
      acc_it++;
      }
    }
  } else if (warp == MMA_WARP) {
    // ===== MMA issuer ==================================================================================================
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0;
      for (int i1 = blockIdx.x; i1 < Mb; i1 += gridDim.x) {
        const int n = __ldg(pos + i1 + 1) - __ldg(pos + i1);
        if (n == 0) continue;
        for (int b0 = 0; b0 < n; b0 += TC::SEG) {                      // one accumulator per segment of <= SEG blocks
        const uint32_t a = acc_it & 1, aph = (acc_it >> 1) & 1;
        mbar_wait(tempty_bar(a), aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * CF::ACC_COLS;
        const int nb = min(TC::SEG, n - b0);
        for (int b = 0; b < nb; b++, it++) {
          const uint32_t rs = it % RS, rph = (it / RS) & 1, ls = it % LS, lph = (it / LS) & 1;
          mbar_wait(raw_full(rs), rph);                              // TMA bytes have landed (async proxy -> async proxy)
          mbar_wait(lo_full(ls), lph);                               // converters have written Bt_lo and [A_hi ; A_lo]
          tc_fence_after();
          const uint32_t bt_hi = raw0 + rs * TC::RAW_BYTES, bt_lo = lo0 + ls * TC::LO_BYTES, ab = bt_lo + TC::BT_BYTES;
#pragma unroll
          for (int g = 0; g < BC / 8; g++) {                         // 8 rows of the B tile = one K step = 1024 bytes per panel
            const uint64_t a_hi = smem_desc_mn32(bt_hi + g * 1024, TC::PANEL_BYTES, 512);
            const uint64_t a_lo = smem_desc_mn32(bt_lo + g * 1024, TC::PANEL_BYTES, 512);
            const uint64_t b_all = smem_desc(ab + g * 2 * CF::AB_LBO, CF::AB_LBO, CF::AB_SBO);
            tc_mma_tf32(d_tmem, a_hi, b_all, TC::idesc(2 * BR), (b | g) != 0);
            tc_mma_tf32(d_tmem, a_lo, b_all, TC::idesc(BR), 1);
          }
          tc_commit(raw_empty(rs));
          tc_commit(lo_empty(ls));
        }
        tc_commit(tfull_bar(a));
        acc_it++;
        }
      }
    }
    __syncwarp();
  } else if (warp == TMA_WARP) {
    // ===== TMA producer: the whole warp reads up to 32 block columns of a row at once, lane 0 issues ===================
    uint32_t it = 0;
    for (int i1 = blockIdx.x; i1 < Mb; i1 += gridDim.x) {
      const int p0 = __ldg(pos + i1), p1 = __ldg(pos + i1 + 1);
      for (int pc = p0; pc < p1; pc += 32) {
        const int mine = pc + lane < p1 ? __ldg(crd + pc + lane) : 0;
        const int cnt = min(32, p1 - pc);
        for (int b = 0; b < cnt; b++, it++) {
          const int j1 = __shfl_sync(0xffffffffu, mine, b);
          if (lane == 0) {
            const uint32_t rs = it % RS, rph = (it / RS) & 1;
            mbar_wait(raw_empty(rs), rph ^ 1);
            const uint32_t st = raw0 + rs * TC::RAW_BYTES;
            mbar_expect_tx(raw_full(rs), TC::RAW_BYTES);
#pragma unroll
            for (int c = 0; c < 4; c++) tma_box_2d(st + c * TC::PANEL_BYTES, &tmB, k0 + 32 * c, j1 * BC, raw_full(rs));
            bulk_copy(st + TC::BT_BYTES, vals + (size_t)(pc + b) * BR * BC, TC::AB_BYTES, raw_full(rs));
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ===== converters: raw ring -> lo ring ================================================================================
    const int ct = threadIdx.x - CONV_WARP0 * 32;                    // 0 .. CW*32-1
    constexpr int CT = CW * 32;
    constexpr int BT_CHUNKS = TC::BT_BYTES / 16, BT_PER = (BT_CHUNKS + CT - 1) / CT;
    constexpr int AB_CHUNKS = TC::AB_BYTES / 16, AB_PER = (AB_CHUNKS + CT - 1) / CT;
    constexpr int JC = BC / 4;                                       // 16-byte chunks per row of the A block
    static_assert(JC == 4 || JC == 8, "diagonal chunk assignment below is written for 16- and 32-wide blocks");
    // chunk index -> (row, chunk in row) along diagonals: the 8 lanes of one shared-memory wavefront read 8 different
    // bank groups of the row-major block ((i2 * JC + jc) % 8) AND write 8 different ones of the K-major operand (i2 % 8);
    // the straight assignment made every store 8-way conflicted (ncu: 246 M conflict wavefronts of 347 M)
    auto ab_row = [](int ch) { return ((ch >> 3) / JC) * 8 + (ch & 7); };
    auto ab_chunk = [](int ch) { return (((ch & 7) * JC >> 3) + (ch >> 3) % JC) % JC; };
    uint32_t it = 0;
    for (int i1 = blockIdx.x; i1 < Mb; i1 += gridDim.x) {
      const int n = __ldg(pos + i1 + 1) - __ldg(pos + i1);
      for (int b = 0; b < n; b++, it++) {
        const uint32_t rs = it % RS, rph = (it / RS) & 1, ls = it % LS, lph = (it / LS) & 1;
        const uint32_t src = raw0 + rs * TC::RAW_BYTES, dst = lo0 + ls * TC::LO_BYTES;
        mbar_wait(raw_full(rs), rph);
        float4 xb[BT_PER], xa[AB_PER];
#pragma unroll
        for (int u = 0; u < BT_PER; u++) {
          const int ch = ct + u * CT;
          if (BT_CHUNKS % CT == 0 || ch < BT_CHUNKS) xb[u] = lds_f4(src + ch * 16);
        }
#pragma unroll
        for (int u = 0; u < AB_PER; u++) {
          const int ch = ct + u * CT;
          if (AB_CHUNKS % CT == 0 || ch < AB_CHUNKS) xa[u] = lds_f4(src + TC::BT_BYTES + (ab_row(ch) * JC + ab_chunk(ch)) * 16);
        }
        mbar_wait(lo_empty(ls), lph ^ 1);
#pragma unroll
        for (int u = 0; u < BT_PER; u++) {
          const int ch = ct + u * CT;
          if (BT_CHUNKS % CT == 0 || ch < BT_CHUNKS) {
            float4 hi, lo;
            split_tf32(xb[u], hi, lo);
            sts_f4(dst + ch * 16, lo);                               // same swizzled position as the raw element
          }
        }
#pragma unroll
        for (int u = 0; u < AB_PER; u++) {
          const int ch = ct + u * CT;
          if (AB_CHUNKS % CT == 0 || ch < AB_CHUNKS) {
            const int i2 = ab_row(ch), jc = ab_chunk(ch);            // row-major A block: row i2, columns 4 jc .. 4 jc + 3
            float4 hi, lo;
            split_tf32(xa[u], hi, lo);
            const uint32_t off = (uint32_t)(jc * (2 * BR / 8) + (i2 >> 3)) * 128 + (uint32_t)(i2 & 7) * 16;
            sts_f4(dst + TC::BT_BYTES + off, hi);
            sts_f4(dst + TC::BT_BYTES + (BR / 8) * 128 + off, lo);
          }
        }
        fence_async_smem();
        mbar_arrive(lo_full(ls));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CF::TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

template <int BR, int BC, int RS, int LS, int CW, int MINB>
static int launch(const int* pos, const int* crd, const float* vals, const float* B, float* C, int Mb, int Nb, int K) {
  using TC = TCfg<BR, BC>;
  EncodeTiledFn enc = encode_tiled();
  if (!enc) return fail(TACO_B200_ERR_CUDA, "bspmm: cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)Nb * BC};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)BC}, estr[2] = {1, 1};
  const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)B, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TACO_B200_ERR_CUDA, "bspmm: cuTensorMapEncodeTiled failed (%d)", (int)r);
  constexpr int threads = (EPI_WARPS + 2 + CW) * 32;
  const int smem = RS * TC::RAW_BYTES + LS * TC::LO_BYTES + 8 * (2 * RS + 2 * LS + 4) + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    TB_CUDA(cudaFuncSetAttribute(bspmm_tma_kernel<BR, BC, RS, LS, CW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int gx = num_sms() * MINB;
  const int gy = (K + TILE_K - 1) / TILE_K;
  gx = (gx + gy - 1) / gy;
  if (gx > Mb) gx = Mb;
  if (gx < 1) gx = 1;
  bspmm_tma_kernel<BR, BC, RS, LS, CW, MINB><<<dim3(gx, gy), threads, smem, stream()>>>(tm, pos, crd, vals, C, Mb, K);
  return TACO_B200_OK;
}

}  // namespace tma

}  // namespace tc

// TACO_B200_BSPMM_TC=0 forces the CUDA-core kernel (reference order), =1 (default) uses tensor cores where they apply.
static bool tc_enabled() {
  static const int v = getenv("TACO_B200_BSPMM_TC") ? atoi(getenv("TACO_B200_BSPMM_TC")) : 1;
  return v != 0;
}

template <typename T>
static int bspmm_launch(const BcsrView& A, const int* pos, const int* crd, const T* vals, const T* B, T* C, int K) {
  if (A.Mb == 0 || K == 0) return TACO_B200_OK;
  ProfScope ps("bspmm_bcsr");
  if constexpr (sizeof(T) == 4) {
    const bool aligned = (K % 4 == 0) && (((uintptr_t)B & 15) == 0) && (((uintptr_t)vals & 15) == 0) &&
                         (long long)A.Nb * A.bc <= INT32_MAX && (K + 127) / 128 <= 65535;   // TMA row coordinates are int32; grid.y

    static const int variant = getenv("TACO_B200_BSPMM_VARIANT") ? atoi(getenv("TACO_B200_BSPMM_VARIANT")) : 0;
    if (tc_enabled() && aligned) {
      // (raw stages, lo stages, converter warps, CTAs per SM); measured 1.59 / 2.00 / 1.93 / 1.60 ms for variants 0, 11, 12,
      // 13 at the bench config (10.2 GB of DRAM traffic -> 6.4 TB/s, the measured HBM peak); 16x16 blocks: 2.86 / 2.79 / 4.62 /
      // 3.39 ms.  The register-staged predecessor of this kernel (global -> registers -> smem, 2.64 ms) is in the history.
      if (A.br == 32 && A.bc == 32) {
        count_launch(1);
        switch (variant) {
          case 11: return tc::tma::launch<32, 32, 4, 2, 8, 1>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
          case 12: return tc::tma::launch<32, 32, 6, 3, 8, 1>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
          case 13: return tc::tma::launch<32, 32, 3, 2, 8, 2>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
          default: return tc::tma::launch<32, 32, 3, 2, 4, 2>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
        }
      }
      if (A.br == 16 && A.bc == 16) {
        count_launch(1);
        switch (variant) {
          case 11: return tc::tma::launch<16, 16, 6, 3, 4, 2>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
          case 12: return tc::tma::launch<16, 16, 8, 4, 8, 1>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
          case 13: return tc::tma::launch<16, 16, 4, 2, 4, 2>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
          default: return tc::tma::launch<16, 16, 6, 3, 8, 2>(pos, crd, vals, B, C, A.Mb, A.Nb, K);
        }
      }
    }
  }
  constexpr int WARPS = 8, RJ = 4;
  const size_t smem = sizeof(T) * (size_t)A.br * A.bc;
  if (smem > 48 * 1024) return fail(TACO_B200_ERR_UNSUPPORTED, "bspmm: blocks larger than 48 KB are not supported");
  dim3 grid(A.Mb, (K + 31) / 32);
  bspmm_rows_kernel<T, WARPS, RJ><<<grid, WARPS * 32, smem, stream()>>>(pos, crd, vals, B, C, A.br, A.bc, K);
  count_launch(1);
  TB_CUDA(cudaGetLastError());
  return TACO_B200_OK;
}

static int bspmm_views(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B, DenseView* Cv, BcsrView* Av, DenseView* Bv) {
  TB_TRY(ensure_init());
  TB_TRY(view_dense(C, 3, "C", Cv));
  TB_TRY(view_bcsr(A, "A", Av));
  TB_TRY(view_dense(B, 3, "B", Bv));
  for (int l = 0; l < 3; l++)
    if (Cv->mode_order[l] != l || Bv->mode_order[l] != l) return fail(TACO_B200_ERR_FORMAT, "bspmm: B and C must be row-major");
  if (Cv->dim[0] != Av->Mb || Cv->dim[1] != Av->br || Bv->dim[0] != Av->Nb || Bv->dim[1] != Av->bc || Cv->dim[2] != Bv->dim[2])
    return fail(TACO_B200_ERR_ARG, "bspmm: dimension mismatch C[%d x %d x %d] = A[%d x %d x %d x %d] * B[%d x %d x %d]", Cv->dim[0],
                Cv->dim[1], Cv->dim[2], Av->Mb, Av->Nb, Av->br, Av->bc, Bv->dim[0], Bv->dim[1], Bv->dim[2]);
  if (Cv->dt != Av->dt || Bv->dt != Av->dt) return fail(TACO_B200_ERR_FORMAT, "bspmm: mixed component types");
  return TACO_B200_OK;
}

static int bspmv_views(taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c, DenseView* av, BcsrView* Av, DenseView* cv) {
  TB_TRY(ensure_init());
  TB_TRY(view_dense(a, 2, "a", av));
  TB_TRY(view_bcsr(A, "A", Av));
  TB_TRY(view_dense(c, 2, "c", cv));
  for (int l = 0; l < 2; l++)
    if (av->mode_order[l] != l || cv->mode_order[l] != l) return fail(TACO_B200_ERR_FORMAT, "bspmv: a and c must be row-major");
  if (av->dim[0] != Av->Mb || av->dim[1] != Av->br || cv->dim[0] != Av->Nb || cv->dim[1] != Av->bc)
    return fail(TACO_B200_ERR_ARG, "bspmv: dimension mismatch a[%d x %d] = A[%d x %d x %d x %d] * c[%d x %d]", av->dim[0], av->dim[1],
                Av->Mb, Av->Nb, Av->br, Av->bc, cv->dim[0], cv->dim[1]);
  if (av->dt != Av->dt || cv->dt != Av->dt) return fail(TACO_B200_ERR_FORMAT, "bspmv: mixed component types");
  return TACO_B200_OK;
}

}  // namespace tb

using namespace tb;

extern "C" {

int taco_b200_bspmm_assemble(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  DenseView Cv, Bv; BcsrView Av;
  TB_TRY(bspmm_views(C, A, B, &Cv, &Av, &Bv));
  void* p = result_alloc(Cv.count() * dsize(Cv.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "bspmm: cannot allocate result");
  C->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

int taco_b200_bspmm_compute(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  DenseView Cv, Bv; BcsrView Av;
  TB_TRY(bspmm_views(C, A, B, &Cv, &Av, &Bv));
  int32_t nnzb = 0;
  TB_TRY(bcsr_nnzb(Av, A->vals_size, &nnzb));
  if (nnzb < 0 || (long long)nnzb * Av.br * Av.bc > (long long)INT32_MAX * 16) return fail(TACO_B200_ERR_ARG, "bspmm: bad block count %d", nnzb);
  if (!Cv.vals) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
  const int K = Bv.dim[2];
  const size_t es = dsize(Av.dt);
  In pos, crd, vals, bin; Out cout;
  TB_TRY(pos.acquire(Av.pos, sizeof(int32_t) * ((size_t)Av.Mb + 1)));
  TB_TRY(crd.acquire(Av.crd ? (void*)Av.crd : (void*)Av.pos, sizeof(int32_t) * (size_t)nnzb));
  TB_TRY(vals.acquire(Av.vals ? Av.vals : (void*)Av.pos, es * (size_t)nnzb * Av.br * Av.bc));
  TB_TRY(bin.acquire(Bv.vals, es * (size_t)Av.Nb * Av.bc * K));
  TB_TRY(cout.acquire(Cv.vals, es * (size_t)Av.Mb * Av.br * K));
  if (Av.dt == DType::F32)
    TB_TRY(bspmm_launch<float>(Av, pos.as<int>(), crd.as<int>(), vals.as<float>(), bin.as<float>(), cout.as<float>(), K));
  else
    TB_TRY(bspmm_launch<double>(Av, pos.as<int>(), crd.as<int>(), vals.as<double>(), bin.as<double>(), cout.as<double>(), K));
  TB_CUDA(cudaGetLastError());
  TB_TRY(cout.commit());
  return finish_call();
}

int taco_b200_bspmm_evaluate(taco_tensor_t* C, taco_tensor_t* A, taco_tensor_t* B) {
  TB_TRY(taco_b200_bspmm_assemble(C, A, B));
  return taco_b200_bspmm_compute(C, A, B);
}

int taco_b200_bspmv_assemble(taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c) {
  DenseView av, cv; BcsrView Av;
  TB_TRY(bspmv_views(a, A, c, &av, &Av, &cv));
  void* p = result_alloc(av.count() * dsize(av.dt));
  if (!p) return fail(TACO_B200_ERR_ALLOC, "bspmv: cannot allocate result");
  a->vals = (uint8_t*)p;
  return TACO_B200_OK;
}

int taco_b200_bspmv_compute(taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c) {
  DenseView av, cv; BcsrView Av;
  TB_TRY(bspmv_views(a, A, c, &av, &Av, &cv));
  int32_t nnzb = 0;
  TB_TRY(bcsr_nnzb(Av, A->vals_size, &nnzb));
  if (nnzb < 0) return fail(TACO_B200_ERR_ARG, "bspmv: bad block count %d", nnzb);
  if (!av.vals) return fail(TACO_B200_ERR_ARG, "NULL result array (call assemble first)");
  const size_t es = dsize(Av.dt);
  In pos, crd, vals, cin; Out aout;
  TB_TRY(pos.acquire(Av.pos, sizeof(int32_t) * ((size_t)Av.Mb + 1)));
  TB_TRY(crd.acquire(Av.crd ? (void*)Av.crd : (void*)Av.pos, sizeof(int32_t) * (size_t)nnzb));
  TB_TRY(vals.acquire(Av.vals ? Av.vals : (void*)Av.pos, es * (size_t)nnzb * Av.br * Av.bc));
  TB_TRY(cin.acquire(cv.vals, es * (size_t)Av.Nb * Av.bc));
  TB_TRY(aout.acquire(av.vals, es * (size_t)Av.Mb * Av.br));
  const long long rows = (long long)Av.Mb * Av.br;
  if (rows > 0) {
    ProfScope ps("bspmv_bcsr");
    static const int variant = getenv("TACO_B200_BSPMV_VARIANT") ? atoi(getenv("TACO_B200_BSPMV_VARIANT")) : 0;
    constexpr int WARPS = 4;
    const size_t smem = (size_t)WARPS * ((size_t)Av.br * (Av.bc + 1) + Av.bc) * es;
    const bool staged = variant == 0 && Av.br <= 32 && smem <= 96 * 1024 && ((size_t)Av.bc * es) % 16 == 0 &&
                        (size_t)Av.br * Av.bc * es <= 16 * 32 * 16 && (((uintptr_t)vals.dptr) & 15) == 0;
    if (staged) {
      long long ctas = ((long long)Av.Mb + WARPS - 1) / WARPS;
      const int g = (int)(ctas < (1 << 20) ? ctas : (1 << 20));
      if (Av.dt == DType::F32) {
        static bool cfg = false;
        if (!cfg) { TB_CUDA(cudaFuncSetAttribute(bspmv_warp_kernel<float, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfg = true; }
        bspmv_warp_kernel<float, WARPS><<<g, WARPS * 32, smem, stream()>>>(pos.as<int>(), crd.as<int>(), vals.as<float>(), cin.as<float>(),
                                                                           aout.as<float>(), Av.Mb, Av.br, Av.bc);
      } else {
        static bool cfg = false;
        if (!cfg) { TB_CUDA(cudaFuncSetAttribute(bspmv_warp_kernel<double, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); cfg = true; }
        bspmv_warp_kernel<double, WARPS><<<g, WARPS * 32, smem, stream()>>>(pos.as<int>(), crd.as<int>(), vals.as<double>(), cin.as<double>(),
                                                                            aout.as<double>(), Av.Mb, Av.br, Av.bc);
      }
      count_launch(1);
      TB_CUDA(cudaGetLastError());
      TB_TRY(aout.commit());
      return finish_call();
    }
    const int grid = (int)((rows + 255) / 256);
    if (Av.dt == DType::F32)
      bspmv_kernel<float><<<grid, 256, 0, stream()>>>(pos.as<int>(), crd.as<int>(), vals.as<float>(), cin.as<float>(), aout.as<float>(),
                                                     Av.Mb, Av.br, Av.bc);
    else
      bspmv_kernel<double><<<grid, 256, 0, stream()>>>(pos.as<int>(), crd.as<int>(), vals.as<double>(), cin.as<double>(),
                                                      aout.as<double>(), Av.Mb, Av.br, Av.bc);
    count_launch(1);
    TB_CUDA(cudaGetLastError());
  }
  TB_TRY(aout.commit());
  return finish_call();
}

int taco_b200_bspmv_evaluate(taco_tensor_t* a, taco_tensor_t* A, taco_tensor_t* c) {
  TB_TRY(taco_b200_bspmv_assemble(a, A, c));
  return taco_b200_bspmv_compute(a, A, c);
}

}  // extern "C"
