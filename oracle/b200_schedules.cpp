// oracle/b200_schedules.cpp -- TEST INFRASTRUCTURE ONLY (built by oracle/Makefile, target ref_b200).
//
// An ordinary taco C++ program, linked against the reference library with patches/module_b200.patch applied
// (oracle/_ref_b200/libtaco.so) and run with TACO_B200=1.  For every GPU statement of the reference's scheduling tests
// (/root/reference/test/tests-scheduling-eval.cpp:1210-1587: spmvGPU, spmmGPU, spmmDCSRGPU, sddmmGPU, ttvGPU, ttmGPU,
// mttkrpGPU) it
//   1. applies the reference's own GPU schedule (the directives and parameters of :193-369) and calls plain
//      compile() / assemble() / compute(): Module::compile() classifies the scheduled statement and binds it to
//      libtaco_b200.so -- no compileSource(), no stub file;
//   2. computes the same statement with the reference's C codegen on the host (CUDA code generation switched off for that
//      tensor), and compares the two results with the reference's equals().
// Also: an unscheduled statement under CUDA code generation (the default GPU schedule of parallelizeOuterLoop), operands
// written in another order (y(i) = x(j) * A(i,j)), the sparse-output statements (SpAdd / SpGEMM, no reference GPU
// schedule exists), and a statement that is GPU-scheduled but not a hot-path pattern (must raise, never run on the CPU).
// Prints one "<name> OK|FAIL" line per case; exit code = number of failures.
#include <cstdlib>
#include <iostream>
#include <string>

#include "taco.h"
#include "taco/cuda.h"
#include "taco/index_notation/transformations.h"

using namespace taco;

static const int WARP = 32;
static IndexVar i("i"), j("j"), k("k"), l("l");
static int failures = 0;

static double val() { return (double)(rand() % 5 + 1); }        // small integers: sums are exact in any order

static void fillSparse(TensorBase& t, double density) {
  std::vector<int> dims = t.getDimensions();
  std::vector<int> c(dims.size(), 0);
  size_t total = 1;
  for (int d : dims) total *= (size_t)d;
  for (size_t q = 0; q < total; q++) {
    size_t r = q;
    for (int m = (int)dims.size() - 1; m >= 0; m--) { c[m] = (int)(r % dims[m]); r /= dims[m]; }
    if ((double)rand() / RAND_MAX < density) t.insert(c, val());
  }
  t.pack();
}

template <typename MakeStmt>
static void check(const std::string& name, TensorBase gpu, TensorBase cpu, MakeStmt schedule) {
  try {
    set_CUDA_codegen_enabled(true);
    IndexStmt stmt = gpu.getAssignment().concretize();
    stmt = schedule(stmt);
    if (stmt.defined()) gpu.compile(stmt); else gpu.compile();
    gpu.assemble();
    gpu.compute();
    const bool bound = gpu.getSource().find("taco_b200_") != std::string::npos;   // the module's source is the forwarding stub
    set_CUDA_codegen_enabled(false);
    cpu.compile();
    cpu.assemble();
    cpu.compute();
    set_CUDA_codegen_enabled(true);
    const bool same = equals(cpu, gpu);
    std::cout << name << (bound && same ? " OK" : " FAIL") << (bound ? "" : " (not bound to libtaco_b200)")
              << (same ? "" : " (differs from the C codegen)") << std::endl;
    if (!(bound && same)) failures++;
  } catch (const TacoException& e) {
    set_CUDA_codegen_enabled(true);
    std::cout << name << " FAIL (exception: " << e.what() << ")" << std::endl;
    failures++;
  }
}

int main() {
  if (!getenv("TACO_B200")) { std::cerr << "run with TACO_B200=1 and TACO_B200_LIB=<path to libtaco_b200.so>" << std::endl; return 2; }
  srand(4357);
  const int NNZ_PER_THREAD = 8, BLOCK = 256;

  {   // ---- spmvGPU (scheduleSpMVGPU :193-209) --------------------------------------------------------------------
    const int n = 1021, m = 1039;
    Tensor<double> A("A", {n, m}, CSR), x("x", {m}, Format({Dense}));
    fillSparse(A, 0.05);
    fillSparse(x, 1.0);
    Tensor<double> y("y", {n}, Format({Dense})), e("e", {n}, Format({Dense}));
    IndexExpr pre = A(i, j) * x(j);
    y(i) = pre;
    e(i) = A(i, j) * x(j);
    check("spmvGPU", y, e, [&](IndexStmt s) {
      IndexVar f("f"), fpos("fpos"), fpos1("fpos1"), fpos2("fpos2"), block("block"), warp("warp"), thread("thread"),
          thread_nz("thread_nz"), thread_nz_pre("thread_nz_pre");
      TensorVar precomputed("precomputed", Type(Float64, {Dimension(thread_nz)}), taco::dense);
      return s.fuse(i, j, f).pos(f, fpos, A(i, j)).split(fpos, block, fpos1, NNZ_PER_THREAD * BLOCK)
          .split(fpos1, warp, fpos2, NNZ_PER_THREAD * WARP).split(fpos2, thread, thread_nz, NNZ_PER_THREAD)
          .reorder({block, warp, thread, thread_nz}).precompute(pre, thread_nz, thread_nz_pre, precomputed)
          .unroll(thread_nz_pre, NNZ_PER_THREAD)
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
    });
    // the same statement without a schedule (CUDA code generation on: parallelizeOuterLoop's default GPU schedule)
    Tensor<double> y2("y2", {n}, Format({Dense})), e2("e2", {n}, Format({Dense}));
    y2(i) = A(i, j) * x(j);
    e2(i) = A(i, j) * x(j);
    check("spmvDefaultSchedule", y2, e2, [&](IndexStmt) { return IndexStmt(); });
    // operands written in the other order: taco packs (y, x, A), the library's kernel takes (y, A, x)
    Tensor<double> y3("y3", {n}, Format({Dense})), e3("e3", {n}, Format({Dense}));
    y3(i) = x(j) * A(i, j);
    e3(i) = x(j) * A(i, j);
    check("spmvCommuted", y3, e3, [&](IndexStmt) { return IndexStmt(); });
  }
  {   // ---- spmmGPU (scheduleSpMMGPU :249-268) and spmmDCSRGPU (scheduleSpMMNZRowsGPU :358-369) -----------------------
    const int n = 1021, m = 1039, K = 128;
    Tensor<double> A("A", {n, m}, CSR), B("B", {m, K}, Format({Dense, Dense}));
    fillSparse(A, 0.03);
    fillSparse(B, 1.0);
    Tensor<double> C("C", {n, K}, Format({{Dense, Dense}, {1, 0}})), E("E", {n, K}, Format({{Dense, Dense}, {1, 0}}));
    C(i, k) = A(i, j) * B(j, k);
    E(i, k) = A(i, j) * B(j, k);
    check("spmmGPU", C, E, [&](IndexStmt s) {
      const int NNZ_PER_WARP = 8;
      IndexVar f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
          dense_val("dense_val"), thread("thread");
      return s.reorder({i, j, k}).fuse(i, j, f).pos(f, fpos, A(i, j)).split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP))
          .split(fpos1, warp, nnz, NNZ_PER_WARP).split(k, dvu, thread, WARP).reorder({block, warp, thread, dvu, nnz})
          .bound(dvu, dense_val, 4, BoundType::MaxExact)
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
    });
    Tensor<double> Ad("Ad", {n, m}, Format({Sparse, Sparse}));
    fillSparse(Ad, 0.01);
    Tensor<double> C2("C2", {n, K}, Format({Dense, Dense})), E2("E2", {n, K}, Format({Dense, Dense}));
    C2(i, k) = Ad(i, j) * B(j, k);
    E2(i, k) = Ad(i, j) * B(j, k);
    check("spmmDCSRGPU", C2, E2, [&](IndexStmt s) {
      const int NZ_ROWS_PER_WARP = 4;
      IndexVar ip("ip"), ip1("ip1"), block("block"), warp("warp"), warp_row("warp_row"), thread("thread"), thread_col("thread_col");
      return s.pos(i, ip, Ad(i, j)).split(ip, block, ip1, NZ_ROWS_PER_WARP * (BLOCK / WARP)).split(ip1, warp, warp_row, NZ_ROWS_PER_WARP)
          .split(k, thread, thread_col, 32).reorder({block, warp, warp_row, thread, thread_col, j})
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
    });
  }
  {   // ---- sddmmGPU (scheduleSDDMMGPU :270-287; dense result, D indexed (contraction, column)) -------------------------
    const int n = 102, m = 103, J = 128;
    Tensor<double> B("B", {n, m}, CSR), C("C", {n, J}, Format({Dense, Dense})), D("D", {J, m}, Format({Dense, Dense}));
    fillSparse(B, 0.3);
    fillSparse(C, 1.0);
    fillSparse(D, 1.0);
    Tensor<double> A("A", {n, m}, Format({Dense, Dense})), E("E", {n, m}, Format({Dense, Dense}));
    A(i, k) = B(i, k) * C(i, j) * D(j, k);
    E(i, k) = B(i, k) * C(i, j) * D(j, k);
    check("sddmmGPU", A, E, [&](IndexStmt s) {
      const int NNZ_PER_WARP = 8 * 32, CO_FACTOR = 4;
      IndexVar f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
          dense_val("dense_val"), thread("thread");
      return s.reorder({i, k, j}).fuse(i, k, f).pos(f, fpos, B(i, k)).split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP))
          .split(fpos1, warp, nnz, NNZ_PER_WARP).split(j, dvu, thread, WARP).bound(dvu, dense_val, CO_FACTOR, BoundType::MaxExact)
          .reorder({block, warp, nnz, thread, dense_val}).unroll(dense_val, CO_FACTOR)
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::Atomics)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::ParallelReduction);
    });
  }
  {   // ---- ttvGPU (:308-325), ttmGPU (:289-306), mttkrpGPU (:327-342) ----------------------------------------------------
    const int I = 102, J = 103, K = 105, R = 32;
    Tensor<double> B("B", {I, J, K}, Format({Sparse, Sparse, Sparse})), c("c", {K}, Format({Dense}));
    fillSparse(B, 0.1);
    fillSparse(c, 1.0);
    Tensor<double> A("A", {I, J}, Format({Dense, Dense})), E("E", {I, J}, Format({Dense, Dense}));
    IndexExpr pre = B(i, j, k) * c(k);
    A(i, j) = pre;
    E(i, j) = B(i, j, k) * c(k);
    check("ttvGPU", A, E, [&](IndexStmt s) {
      const int NNZ_PER_WARP = 8 * 32;
      IndexVar jk("jk"), f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), fpos2("fpos2"), thread("thread"),
          thread_nz("thread_nz"), thread_nz_pre("thread_nz_pre");
      TensorVar precomputed("precomputed", Type(Float64, {Dimension(thread_nz)}), taco::dense);
      return s.fuse(j, k, jk).fuse(i, jk, f).pos(f, fpos, B(i, j, k)).split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP))
          .split(fpos1, warp, fpos2, NNZ_PER_WARP).split(fpos2, thread, thread_nz, NNZ_PER_WARP / WARP)
          .reorder({block, warp, thread, thread_nz}).precompute(pre, thread_nz, thread_nz_pre, precomputed)
          .unroll(thread_nz_pre, NNZ_PER_WARP / WARP)
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
    });
    Tensor<double> Cm("Cm", {K, R}, Format({Dense, Dense}));
    fillSparse(Cm, 1.0);
    Tensor<double> A3("A3", {I, J, R}, Format({Dense, Dense, Dense})), E3("E3", {I, J, R}, Format({Dense, Dense, Dense}));
    A3(i, j, l) = B(i, j, k) * Cm(k, l);
    E3(i, j, l) = B(i, j, k) * Cm(k, l);
    check("ttmGPU", A3, E3, [&](IndexStmt s) {
      const int NNZ_PER_WARP = 8 * 32, CO_FACTOR = 1;
      IndexVar jk("jk"), f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
          dense_val("dense_val"), thread("thread");
      return s.reorder({i, j, k, l}).fuse(j, k, jk).fuse(i, jk, f).pos(f, fpos, B(i, j, k))
          .split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP)).split(fpos1, warp, nnz, NNZ_PER_WARP)
          .split(l, dvu, thread, WARP).bound(dvu, dense_val, CO_FACTOR, BoundType::MaxExact)
          .reorder({block, warp, nnz, thread, dense_val}).unroll(dense_val, CO_FACTOR)
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
    });
    Tensor<double> Cf("Cf", {J, R}, Format({Dense, Dense})), Df("Df", {K, R}, Format({Dense, Dense}));
    fillSparse(Cf, 1.0);
    fillSparse(Df, 1.0);
    Tensor<double> Am("Am", {I, R}, Format({Dense, Dense})), Em("Em", {I, R}, Format({Dense, Dense}));
    Am(i, j) = B(i, k, l) * Cf(k, j) * Df(l, j);
    Em(i, j) = B(i, k, l) * Cf(k, j) * Df(l, j);
    check("mttkrpGPU", Am, Em, [&](IndexStmt s) {
      const int NNZ_PER_WARP = 16;
      IndexVar kl("kl"), f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
          dense_val("dense_val"), thread("thread");
      return s.reorder({i, k, l, j}).fuse(k, l, kl).fuse(i, kl, f).pos(f, fpos, B(i, k, l))
          .split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP)).split(fpos1, warp, nnz, NNZ_PER_WARP)
          .split(j, dvu, thread, WARP).bound(dvu, dense_val, 1, BoundType::MaxExact).reorder({block, warp, dense_val, thread, nnz})
          .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
          .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
          .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
    });
  }
  {   // ---- sparse outputs: SpAdd and SpGEMM with CSR results (GPU assembly; the reference has no GPU schedule for them) ----
    const int n = 300;
    Tensor<double> A("A", {n, n}, CSR), B("B", {n, n}, CSR);
    fillSparse(A, 0.04);
    fillSparse(B, 0.04);
    Tensor<double> C("C", {n, n}, CSR), E("E", {n, n}, CSR);
    C(i, j) = A(i, j) + B(i, j);
    E(i, j) = A(i, j) + B(i, j);
    check("spaddCSR", C, E, [&](IndexStmt) { return IndexStmt(); });
    Tensor<double> G("G", {n, n}, CSR), H("H", {n, n}, CSR);
    G(i, k) = A(i, j) * B(j, k);
    H(i, k) = A(i, j) * B(j, k);
    check("spgemmCSR", G, H, [&](IndexStmt) { return IndexStmt(); });
  }
  {   // ---- GPU-scheduled but not a hot-path pattern: must raise (the GPU path has no CPU fallback) ---------------------------
    const int n = 64;
    Tensor<double> A("A", {n, n}, CSR), z("z", {n}, Format({Dense})), x("x", {n}, Format({Dense}));
    fillSparse(A, 0.1);
    fillSparse(x, 1.0);
    z(i) = A(i, j) * x(j) * x(i);
    bool raised = false;
    try {
      set_CUDA_codegen_enabled(true);
      IndexVar block("block"), thread("thread");
      IndexStmt s = z.getAssignment().concretize();
      s = s.split(i, block, thread, 32).parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
              .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::NoRaces);
      z.compile(s);
    } catch (const TacoException& e) {
      raised = std::string(e.what()).find("not a hot-path pattern") != std::string::npos;
    }
    std::cout << "offPathStatementRefused" << (raised ? " OK" : " FAIL") << std::endl;
    if (!raised) failures++;
  }
  std::cout << (failures ? "FAILED " : "ALL OK ") << failures << std::endl;
  return failures;
}
