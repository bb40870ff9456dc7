// tools/tc_probe.cu -- bring-up probe for the tcgen05 path of bcsr.cu: ONE tcgen05.mma (kind::tf32, M=128, N, K=8) on
// operand images the HOST lays out under a stated hypothesis about the no-swizzle canonical layouts, accumulator read
// back with tcgen05.ld and compared with the exact product.  Prints which (layout, LBO/SBO assignment) hypotheses hold.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tc_probe tools/tc_probe.cu && gpurun_out/tc_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout = 0) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

struct Params { uint32_t lboA, sboA, lboB, sboB, idesc, n, fence, layoutA; };

__global__ void __launch_bounds__(128) probe(const float* imgA, const float* imgB, Params P, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  unsigned char* gen = smem + (base - (uint32_t)__cvta_generic_to_shared(smem));
  float* sA = (float*)gen;                 // 16 KB
  float* sB = (float*)(gen + 16384);       // 16 KB
  const uint32_t bar = base + 32768, slot = base + 32768 + 16;
  for (int i = threadIdx.x; i < 4096; i += 128) { sA[i] = imgA[i]; sB[i] = imgB[i]; }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (P.fence) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot) : "memory");
  if (threadIdx.x == 0) {
    const uint64_t da = smem_desc(base, P.lboA, P.sboA, P.layoutA), db = smem_desc(base + 16384, P.lboB, P.sboB);
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                 "l"(da), "l"(db), "r"(P.idesc), "r"(0u) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D1;\nbra W1;\nD1:\n}" ::"r"(bar) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t c = 0; c < P.n; c += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; j++) out[(warp * 32 + lane) * P.n + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

static uint32_t idesc(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
}

int main() {
  const int M = 128, K = 8;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, 16384); cudaMalloc(&dB, 16384); cudaMalloc(&dO, 128 * 64 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int N : {16, 32}) {
    std::vector<float> A(M * K), B(N * K), want(M * N);
    for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) A[m * K + k] = (float)((m * 3 + k * 5) % 11 + 1);
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) B[n * K + k] = (float)((n * 7 + k * 2) % 13 - 6);
    for (int m = 0; m < M; m++) for (int n = 0; n < N; n++) { float s = 0; for (int k = 0; k < K; k++) s += A[m * K + k] * B[n * K + k]; want[m * N + n] = s; }
    // operand B image, K-major: core matrix (8 n x 4 k) = 128 B at (kc * (N/8) + ng) * 128
    std::vector<float> imgB(4096, 0.f);
    for (int n = 0; n < N; n++) for (int k = 0; k < K; k++) imgB[((k / 4) * (N / 8) + n / 8) * 32 + (n % 8) * 4 + (k % 4)] = B[n * K + k];
    const uint32_t lboB = (N / 8) * 128, sboB = 128;
    for (int hyp = 0; hyp < 16; hyp++) {
      std::vector<float> imgA(4096, 0.f);
      Params P{};
      P.n = N; P.fence = 1; P.lboB = lboB; P.sboB = sboB;
      const char* name = "";
      if (hyp == 0 || hyp == 1 || hyp == 4) {       // A K-major: core (8 m x 4 k) at (kc * 16 + mg) * 128
        for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) imgA[((k / 4) * 16 + m / 8) * 32 + (m % 8) * 4 + (k % 4)] = A[m * K + k];
        P.idesc = idesc(N, 0, 0);
        if (hyp == 0) { P.lboA = 2048; P.sboA = 128; name = "A K-major  LBO=k-chunk stride, SBO=row-group stride        "; }
        if (hyp == 1) { P.lboA = 128; P.sboA = 2048; name = "A K-major  LBO/SBO swapped (A only)                        "; }
        if (hyp == 4) { P.lboA = 128; P.sboA = 2048; P.lboB = sboB; P.sboB = lboB; name = "A,B K-major both LBO/SBO swapped                           "; }
      } else if (hyp >= 12) {                       // truncation vs rounding of fp32 -> tf32 operands (K-major, as hyp 0)
        // A[m][0] = 1 + 2^-11 + 2^-12 (above half a tf32 ulp), B[n][0] = 1, everything else 0:
        // D = 1 if the tensor core truncates the low 13 bits, 1 + 2^-10 if it rounds to nearest
        if (hyp > 12) continue;
        std::vector<float> imgB2(4096, 0.f);
        for (int n = 0; n < N; n++) imgB2[(n / 8) * 32 + (n % 8) * 4] = 1.0f;
        for (int m = 0; m < M; m++) imgA[(m / 8) * 32 + (m % 8) * 4] = 1.0f + 0.00048828125f + 0.000244140625f;
        P.idesc = idesc(N, 0, 0); P.lboA = 2048; P.sboA = 128;
        cudaMemcpy(dA, imgA.data(), 16384, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, imgB2.data(), 16384, cudaMemcpyHostToDevice);
        probe<<<1, 128, 40000>>>(dA, dB, P, dO);
        cudaDeviceSynchronize();
        float g[2];
        cudaMemcpy(g, dO, 8, cudaMemcpyDeviceToHost);
        printf("N=%2d hyp 12 fp32 -> tf32 operand conversion: D = %.10f (1.0 = truncation, 1.0009765625 = round to nearest)\n", N, g[0]);
        continue;
      } else if (hyp >= 10) {                       // A MN-major, 128-byte swizzle with 32-byte atoms (the only MN-major tf32
        // layout CUTLASS uses: Swizzle<2,5,2> on byte addresses): panel = 32 m (128 B) x 8 k rows, 32-byte chunk ^= k % 4
        const int pbytes = 1024;
        for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) {
          const int mi = m % 32;
          imgA[((m / 32) * pbytes + k * 128 + (((mi / 8) ^ (k % 4)) * 32) + (mi % 8) * 4) / 4] = A[m * K + k];
        }
        P.idesc = idesc(N, 1, 0);
        P.layoutA = 1;
        if (hyp == 10) { P.lboA = 1024; P.sboA = 512; name = "A MN-major SW128/32B-atom LBO=panel(1024), SBO=4-row group(512)"; }
        if (hyp == 11) { P.lboA = 512; P.sboA = 1024; name = "A MN-major SW128/32B-atom LBO=4-row group(512), SBO=panel(1024)"; }
      } else if (hyp >= 6) {                        // A MN-major, 128-byte swizzle: panels of 32 m (128 B) x 8 k rows,
        const int pstride = 2048 / 4;               // 16-byte chunk index XOR (k % 8); panels 2048 B apart
        for (int m = 0; m < M; m++) for (int k = 0; k < K; k++)
          imgA[(m / 32) * pstride + (k % 8) * 32 + ((((m % 32) / 4) ^ (k % 8)) * 4) + (m % 4)] = A[m * K + k];
        P.idesc = idesc(N, 1, 0);
        P.layoutA = 2;
        if (hyp == 6) { P.lboA = 2048; P.sboA = 1024; name = "A MN-major SW128 LBO=panel stride(2048), SBO=k-group(1024)  "; }
        if (hyp == 7) { P.lboA = 1024; P.sboA = 2048; name = "A MN-major SW128 SBO=panel stride(2048), LBO=k-group(1024)  "; }
        if (hyp == 8) { P.lboA = 2048; P.sboA = 2048; name = "A MN-major SW128 LBO=SBO=panel stride(2048)                 "; }
        if (hyp == 9) { P.lboA = 1; P.sboA = 2048; name = "A MN-major SW128 LBO=1(unused), SBO=panel stride(2048)      "; }
      } else {                                      // A MN-major: core (8 k x 4 m) at (m/4) * 128 + (k % 8) * 16
        for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) imgA[(m / 4) * 32 + (k % 8) * 4 + (m % 4)] = A[m * K + k];
        P.idesc = idesc(N, 1, 0);
        if (hyp == 2) { P.lboA = 4096; P.sboA = 128; name = "A MN-major SBO=m-chunk stride(128), LBO=k-group stride(4096)"; }
        if (hyp == 3) { P.lboA = 128; P.sboA = 4096; name = "A MN-major LBO=m-chunk stride(128), SBO=k-group stride(4096)"; }
        if (hyp == 5) { P.lboA = 4096; P.sboA = 128; P.fence = 0; name = "A MN-major as hyp 2 but WITHOUT fence.proxy.async           "; }
      }
      cudaMemcpy(dA, imgA.data(), 16384, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, imgB.data(), 16384, cudaMemcpyHostToDevice);
      cudaMemset(dO, 0xFF, 128 * 64 * 4);
      probe<<<1, 128, 40000>>>(dA, dB, P, dO);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> got(M * N);
      cudaMemcpy(got.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
      int bad = 0, zeros = 0;
      for (int i = 0; i < M * N; i++) { bad += got[i] != want[i]; zeros += got[i] == 0.f; }
      printf("N=%2d hyp %d %s : %s  mismatches %d / %d, zeros %d  [%s]  got %g %g %g %g want %g %g %g %g\n", N, hyp, name,
             bad ? "FAIL" : "PASS", bad, M * N, zeros, cudaGetErrorString(e), got[0], got[1], got[N], got[5 * N + 3], want[0], want[1],
             want[N], want[5 * N + 3]);
    }
  }
  return 0;
}
