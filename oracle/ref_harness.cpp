// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference (tensor-compiler/taco, linked from oracle/_ref/libtaco.so) through its own
// public C++ API, exactly the way the reference's tests do (test/tests-scheduling-eval.cpp):
//     Tensor<T> result; result(i..) = expr; stmt = result.getAssignment().concretize(); [schedule]
//     result.compile(stmt); result.assemble(); result.compute();
// The reference lowers the statement, emits C (src/codegen/codegen_c.cpp), shells out to `cc`
// (src/codegen/module.cpp:111-167) and calls the generated kernel through taco_tensor_t.  Operands are attached
// zero-copy (makeCSR, include/taco/tensor.h:774-797; Index/ModeIndex for CSF, include/taco/storage/index.h:20-75).
//
// usage: taco_ref_harness <kernel> <in.tbin> <out.tbin> [--dtype f64|f32] [--schedule default|cpu]
//                         [--threads N] [--reps R] [--no-out 1]
//   kernel in {spmv, spmm, spmm_dcsr, sddmm, sddmm_dense, mttkrp, spadd, spgemm, ttv, ttm, bspmv, bspmm, pack_csr, pack_dcsr, pack_csf3}
// Prints one JSON line: {"kernel":..., "assemble_ms":[...], "compute_ms":[...], "compile_ms":..., "threads":N}
//
// Input arrays (tbin.h): dims (int32), and per kernel
//   spmv   : A_pos A_crd A_vals x                          -> y
//   spmm   : A_pos A_crd A_vals B                          -> C            dims = n m K
//   sddmm  : B_pos B_crd B_vals C D                        -> A_pos A_crd A_vals      dims = n m K
//   mttkrp : B1_pos B1_crd B2_pos B2_crd B3_pos B3_crd B_vals C D -> A     dims = I K L R
//   ttv    : B1..B3, B_vals, c                             -> A (I x K dense)         dims = I K L
//   ttm    : B1..B3, B_vals, C (L x R)                     -> A (I x K x R dense)     dims = I K L R
//   spadd  : A_* B_*                                       -> C_pos C_crd C_vals      dims = n m
//   spgemm : A_* B_*                                       -> C_pos C_crd C_vals      dims = n m o
//   bspmv  : A_pos A_crd A_vals (blocks br x bc) c (Nb x bc)   -> a (Mb x br)         dims = Mb Nb br bc
//   bspmm  : A_pos A_crd A_vals (blocks br x bc) B (Nb x bc x K) -> C (Mb x br x K)   dims = Mb Nb br bc K
#include <chrono>
#include <iostream>
#include <string>
#include <vector>
#include <cstring>

#include "taco.h"
#include "taco/index_notation/transformations.h"
#include "taco/index_notation/index_notation.h"
#include "taco/storage/index.h"
#include "taco/storage/array.h"
#include "taco/lower/lower.h"
#include "taco/codegen/module.h"
#include "taco/cuda.h"
#include "tbin.h"

using namespace taco;

static double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T> static uint32_t tbin_dtype();
template <> uint32_t tbin_dtype<float>() { return 1; }
template <> uint32_t tbin_dtype<double>() { return 2; }

static tbin_array* need(tbin_file& f, const char* name) {
  tbin_array* a = tbin_get(&f, name);
  if (!a) { std::cerr << "missing input array " << name << std::endl; exit(2); }
  return a;
}

template <typename T>
static Tensor<T> attachDense(const std::string& name, std::vector<int> dims, T* data) {
  std::vector<ModeFormatPack> mf(dims.size(), Dense);
  Tensor<T> t(name, dims, Format(mf));
  auto st = t.getStorage();
  std::vector<ModeIndex> mi;
  size_t n = 1;
  for (int d : dims) { mi.push_back(ModeIndex({makeArray(std::vector<int>{d})})); n *= (size_t)d; }
  st.setIndex(Index(t.getFormat(), mi));
  st.setValues(makeArray(data, n, Array::UserOwns));
  t.setStorage(st);
  return t;
}

template <typename T>
static Tensor<T> attachCSR(const std::string& name, std::vector<int> dims, tbin_file& f, const std::string& p) {
  int* pos = (int*)need(f, (p + "_pos").c_str())->data;
  int* crd = (int*)need(f, (p + "_crd").c_str())->data;
  T* vals = (T*)need(f, (p + "_vals").c_str())->data;
  return makeCSR<T>(name, dims, pos, crd, vals);
}

// Doubly compressed rows {Sparse, Sparse} (the operand of the reference's spmmDCSRGPU test,
// test/tests-scheduling-eval.cpp:1309-1358): level arrays A1_pos[2], A1_crd[nzrows], A2_pos[nzrows+1], A2_crd[nnz]
template <typename T>
static Tensor<T> attachDCSR(const std::string& name, std::vector<int> dims, tbin_file& f, const std::string& p) {
  Tensor<T> t(name, dims, Format({Sparse, Sparse}));
  auto st = t.getStorage();
  std::vector<ModeIndex> mi;
  for (int l = 1; l <= 2; l++) {
    tbin_array* pos = need(f, (p + std::to_string(l) + "_pos").c_str());
    tbin_array* crd = need(f, (p + std::to_string(l) + "_crd").c_str());
    mi.push_back(ModeIndex({makeArray((int*)pos->data, pos->count, Array::UserOwns),
                            makeArray((int*)crd->data, crd->count, Array::UserOwns)}));
  }
  tbin_array* v = need(f, (p + "_vals").c_str());
  st.setIndex(Index(t.getFormat(), mi));
  st.setValues(makeArray((T*)v->data, v->count, Array::UserOwns));
  t.setStorage(st);
  return t;
}

template <typename T>
static Tensor<T> attachCSF3(const std::string& name, std::vector<int> dims, tbin_file& f, const std::string& p) {
  Tensor<T> t(name, dims, Format({Sparse, Sparse, Sparse}));
  auto st = t.getStorage();
  std::vector<ModeIndex> mi;
  for (int l = 1; l <= 3; l++) {
    tbin_array* pos = need(f, (p + std::to_string(l) + "_pos").c_str());
    tbin_array* crd = need(f, (p + std::to_string(l) + "_crd").c_str());
    mi.push_back(ModeIndex({makeArray((int*)pos->data, pos->count, Array::UserOwns),
                            makeArray((int*)crd->data, crd->count, Array::UserOwns)}));
  }
  tbin_array* v = need(f, (p + "_vals").c_str());
  st.setIndex(Index(t.getFormat(), mi));
  st.setValues(makeArray((T*)v->data, v->count, Array::UserOwns));
  t.setStorage(st);
  return t;
}

// Blocked CSR as the reference's own `bspmv` test declares it (test/tests-expr_storage.cpp:939-960):
// order-4 tensor {Dense, Compressed, Dense, Dense} over (block row, block column, row in block, column in block).
template <typename T>
static Tensor<T> attachBCSR(const std::string& name, std::vector<int> dims, tbin_file& f, const std::string& p) {
  Tensor<T> t(name, dims, Format({Dense, Sparse, Dense, Dense}));
  auto st = t.getStorage();
  tbin_array* pos = need(f, (p + "_pos").c_str());
  tbin_array* crd = need(f, (p + "_crd").c_str());
  tbin_array* v = need(f, (p + "_vals").c_str());
  std::vector<ModeIndex> mi;
  mi.push_back(ModeIndex({makeArray(std::vector<int>{dims[0]})}));
  mi.push_back(ModeIndex({makeArray((int*)pos->data, pos->count, Array::UserOwns),
                          makeArray((int*)crd->data, crd->count, Array::UserOwns)}));
  mi.push_back(ModeIndex({makeArray(std::vector<int>{dims[2]})}));
  mi.push_back(ModeIndex({makeArray(std::vector<int>{dims[3]})}));
  st.setIndex(Index(t.getFormat(), mi));
  st.setValues(makeArray((T*)v->data, v->count, Array::UserOwns));
  t.setStorage(st);
  return t;
}

static void dumpSource(const TensorBase& t) {
  static bool done = false;
  if (!done && getenv("TACO_REF_DUMP")) { std::cerr << t.getSource() << std::endl; done = true; }
}

struct Times { std::vector<double> assemble, compute; double compile = 0; };
static bool g_no_out = false;   // --no-out 1: timing runs skip writing the result file

// Schedules follow the reference's own CPU schedules, test/tests-scheduling-eval.cpp:41-184.
template <typename T>
static int run(const std::string& kernel, tbin_file& in, const char* outPath, const std::string& schedule,
               int reps, Times& tm) {
  int* dims = (int*)need(in, "dims")->data;
  IndexVar i("i"), j("j"), k("k"), l("l");
  std::vector<tbin_array> outs;
  bool tuned = (schedule == "cpu");

  for (int rep = 0; rep < reps; rep++) {
    bool last = (rep == reps - 1) && !g_no_out;
    outs.clear();
    if (kernel == "spmv") {
      int n = dims[0], m = dims[1];
      Tensor<T> A = attachCSR<T>("A", {n, m}, in, "A");
      Tensor<T> x = attachDense<T>("x", {m}, (T*)need(in, "x")->data);
      Tensor<T> y("y", {n}, Format({Dense}));
      y(i) = A(i, j) * x(j);
      IndexStmt stmt = y.getAssignment().concretize();
      if (tuned) {
        IndexVar i0("i0"), i1("i1");
        stmt = stmt.split(i, i0, i1, 16).reorder({i0, i1, j})
                   .parallelize(i0, ParallelUnit::CPUThread, OutputRaceStrategy::NoRaces);
      }
      double t0 = now_ms(); if (tuned) y.compile(stmt); else y.compile();
      double t1 = now_ms(); dumpSource(y); y.assemble();
      double t2 = now_ms(); y.compute();
      double t3 = now_ms();
      if (rep == 0) tm.compile = t1 - t0;
      tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
      if (last) {
        tbin_array a; strcpy(a.name, "y"); a.dtype = tbin_dtype<T>(); a.count = n;
        a.data = y.getStorage().getValues().getData(); outs.push_back(a);
        tbin_write(outPath, outs.data(), outs.size());
      }
    } else if (kernel == "spmm" || kernel == "spmm_dcsr") {
      int n = dims[0], m = dims[1], K = dims[2];
      Tensor<T> A = kernel == "spmm" ? attachCSR<T>("A", {n, m}, in, "A") : attachDCSR<T>("A", {n, m}, in, "A");
      Tensor<T> B = attachDense<T>("B", {m, K}, (T*)need(in, "B")->data);
      Tensor<T> C("C", {n, K}, Format({Dense, Dense}));
      C(i, k) = A(i, j) * B(j, k);
      IndexStmt stmt = C.getAssignment().concretize();
      if (tuned && kernel == "spmm") {
        IndexVar i0("i0"), i1("i1"), jpos("jpos"), jpos0("jpos0"), jpos1("jpos1");
        stmt = stmt.split(i, i0, i1, 16).pos(j, jpos, A(i, j)).split(jpos, jpos0, jpos1, 8)
                   .reorder({i0, i1, jpos0, k, jpos1})
                   .parallelize(i0, ParallelUnit::CPUThread, OutputRaceStrategy::NoRaces)
                   .parallelize(k, ParallelUnit::CPUVector, OutputRaceStrategy::IgnoreRaces);
      }
      double t0 = now_ms(); if (tuned) C.compile(stmt); else C.compile();
      double t1 = now_ms(); dumpSource(C); C.assemble();
      double t2 = now_ms(); C.compute();
      double t3 = now_ms();
      if (rep == 0) tm.compile = t1 - t0;
      tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
      if (last) {
        tbin_array a; strcpy(a.name, "C"); a.dtype = tbin_dtype<T>(); a.count = (uint64_t)n * K;
        a.data = C.getStorage().getValues().getData(); outs.push_back(a);
        tbin_write(outPath, outs.data(), outs.size());
      }
    } else if (kernel == "sddmm") {
      int n = dims[0], m = dims[1], K = dims[2];
      Tensor<T> B = attachCSR<T>("B", {n, m}, in, "B");
      Tensor<T> C = attachDense<T>("C", {n, K}, (T*)need(in, "C")->data);
      Tensor<T> D = attachDense<T>("D", {m, K}, (T*)need(in, "D")->data);
      Tensor<T> A("A", {n, m}, CSR);
      A(i, j) = B(i, j) * C(i, k) * D(j, k);
      double t0 = now_ms(); A.compile();
      double t1 = now_ms(); dumpSource(A); A.assemble();
      double t2 = now_ms(); A.compute();
      double t3 = now_ms();
      if (rep == 0) tm.compile = t1 - t0;
      tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
      if (last) {
        int *pos, *crd; T* vals;
        getCSRArrays<T>(A, &pos, &crd, &vals);
        tbin_array a; strcpy(a.name, "A_pos"); a.dtype = 0; a.count = n + 1; a.data = pos; outs.push_back(a);
        strcpy(a.name, "A_crd"); a.dtype = 0; a.count = pos[n]; a.data = crd; outs.push_back(a);
        strcpy(a.name, "A_vals"); a.dtype = tbin_dtype<T>(); a.count = pos[n]; a.data = vals; outs.push_back(a);
        tbin_write(outPath, outs.data(), outs.size());
      }
    } else if (kernel == "sddmm_dense") {
      // the statement of the reference's sddmmGPU test (test/tests-scheduling-eval.cpp:1360-1418): dense result, D indexed
      // (contraction, column):  A(i,k) = B(i,k) * C(i,j) * D(j,k);  dims = I K J
      int n = dims[0], m = dims[1], J = dims[2];
      Tensor<T> B = attachCSR<T>("B", {n, m}, in, "B");
      Tensor<T> C = attachDense<T>("C", {n, J}, (T*)need(in, "C")->data);
      Tensor<T> D = attachDense<T>("D", {J, m}, (T*)need(in, "D")->data);
      Tensor<T> A("A", {n, m}, Format({Dense, Dense}));
      A(i, k) = B(i, k) * C(i, j) * D(j, k);
      double t0 = now_ms(); A.compile();
      double t1 = now_ms(); dumpSource(A); A.assemble();
      double t2 = now_ms(); A.compute();
      double t3 = now_ms();
      if (rep == 0) tm.compile = t1 - t0;
      tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
      if (last) {
        tbin_array a; strcpy(a.name, "A"); a.dtype = tbin_dtype<T>(); a.count = (uint64_t)n * m;
        a.data = A.getStorage().getValues().getData(); outs.push_back(a);
        tbin_write(outPath, outs.data(), outs.size());
      }
    } else if (kernel == "mttkrp" || kernel == "ttv" || kernel == "ttm") {
      int I = dims[0], K = dims[1], L = dims[2];
      Tensor<T> B = attachCSF3<T>("B", {I, K, L}, in, "B");
      if (kernel == "mttkrp") {
        int R = dims[3];
        Tensor<T> C = attachDense<T>("C", {K, R}, (T*)need(in, "C")->data);
        Tensor<T> D = attachDense<T>("D", {L, R}, (T*)need(in, "D")->data);
        Tensor<T> A("A", {I, R}, Format({Dense, Dense}));
        A(i, j) = B(i, k, l) * C(k, j) * D(l, j);
        IndexStmt stmt = A.getAssignment().concretize();
        if (tuned) {
          IndexVar i1("i1"), i2("i2");
          IndexExpr pre = stmt.as<Forall>().getStmt().as<Forall>().getStmt().as<Forall>().getStmt()
                              .as<Forall>().getStmt().as<Assignment>().getRhs().as<Mul>().getA();
          TensorVar w("w", Type(type<T>(), {(size_t)R}), taco::dense);
          stmt = stmt.split(i, i1, i2, 16).reorder({i1, i2, k, l, j});
          stmt = stmt.precompute(pre, j, j, w);
          stmt = stmt.parallelize(i1, ParallelUnit::CPUThread, OutputRaceStrategy::NoRaces);
        }
        double t0 = now_ms(); if (tuned) A.compile(stmt); else A.compile();
        double t1 = now_ms(); dumpSource(A); A.assemble();
        double t2 = now_ms(); A.compute();
        double t3 = now_ms();
        if (rep == 0) tm.compile = t1 - t0;
        tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
        if (last) {
          tbin_array a; strcpy(a.name, "A"); a.dtype = tbin_dtype<T>(); a.count = (uint64_t)I * R;
          a.data = A.getStorage().getValues().getData(); outs.push_back(a);
          tbin_write(outPath, outs.data(), outs.size());
        }
      } else if (kernel == "ttv") {
        Tensor<T> c = attachDense<T>("c", {L}, (T*)need(in, "c")->data);
        Tensor<T> A("A", {I, K}, Format({Dense, Dense}));
        A(i, j) = B(i, j, k) * c(k);
        double t0 = now_ms(); A.compile();
        double t1 = now_ms(); dumpSource(A); A.assemble();
        double t2 = now_ms(); A.compute();
        double t3 = now_ms();
        if (rep == 0) tm.compile = t1 - t0;
        tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
        if (last) {
          tbin_array a; strcpy(a.name, "A"); a.dtype = tbin_dtype<T>(); a.count = (uint64_t)I * K;
          a.data = A.getStorage().getValues().getData(); outs.push_back(a);
          tbin_write(outPath, outs.data(), outs.size());
        }
      } else {
        int R = dims[3];
        Tensor<T> C = attachDense<T>("C", {L, R}, (T*)need(in, "C")->data);
        Tensor<T> A("A", {I, K, R}, Format({Dense, Dense, Dense}));
        A(i, j, l) = B(i, j, k) * C(k, l);
        double t0 = now_ms(); A.compile();
        double t1 = now_ms(); dumpSource(A); A.assemble();
        double t2 = now_ms(); A.compute();
        double t3 = now_ms();
        if (rep == 0) tm.compile = t1 - t0;
        tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
        if (last) {
          tbin_array a; strcpy(a.name, "A"); a.dtype = tbin_dtype<T>(); a.count = (uint64_t)I * K * R;
          a.data = A.getStorage().getValues().getData(); outs.push_back(a);
          tbin_write(outPath, outs.data(), outs.size());
        }
      }
    } else if (kernel == "spadd" || kernel == "spgemm") {
      bool add = (kernel == "spadd");
      int n = dims[0], m = dims[1], o = add ? dims[1] : dims[2];
      Tensor<T> A = attachCSR<T>("A", {n, m}, in, "A");
      Tensor<T> B = attachCSR<T>("B", add ? std::vector<int>{n, m} : std::vector<int>{m, o}, in, "B");
      Tensor<T> C("C", {n, o}, CSR);
      if (add) C(i, j) = A(i, j) + B(i, j);
      else     C(i, k) = A(i, j) * B(j, k);
      IndexStmt stmt = C.getAssignment().concretize();
      if (tuned) {
        // scheduleSpAddCPU / scheduleSpGEMMCPU(doPrecompute=true): two-phase Insert assembly, rows in parallel.
        Assignment assign = add ? stmt.as<Forall>().getStmt().as<Forall>().getStmt().as<Assignment>()
                                : stmt.as<Forall>().getStmt().as<Forall>().getStmt().as<Forall>().getStmt().as<Assignment>();
        TensorVar result = assign.getLhs().getTensorVar();
        stmt = reorderLoopsTopologically(stmt);
        if (!add) {
          IndexVar jj = assign.getLhs().getIndexVars()[1];
          TensorVar w("w", Type(result.getType().getDataType(), {result.getType().getShape().getDimension(1)}), taco::dense);
          stmt = stmt.precompute(assign.getRhs(), jj, jj, w);
        }
        stmt = stmt.assemble(result, AssembleStrategy::Insert, true);
        IndexStmt q = stmt.as<Assemble>().getQueries();
        IndexVar qi = isa<Where>(q) ? q.as<Where>().getConsumer().as<Forall>().getIndexVar()
                                    : q.as<Forall>().getIndexVar();
        stmt = stmt.parallelize(i, ParallelUnit::CPUThread, OutputRaceStrategy::NoRaces)
                   .parallelize(qi, ParallelUnit::CPUThread, OutputRaceStrategy::NoRaces);
      }
      double t0 = now_ms(); if (tuned) C.compile(stmt); else C.compile();
      double t1 = now_ms(); dumpSource(C); C.assemble();
      double t2 = now_ms(); C.compute();
      double t3 = now_ms();
      if (rep == 0) tm.compile = t1 - t0;
      tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
      if (last) {
        int *pos, *crd; T* vals;
        getCSRArrays<T>(C, &pos, &crd, &vals);
        tbin_array a; strcpy(a.name, "C_pos"); a.dtype = 0; a.count = n + 1; a.data = pos; outs.push_back(a);
        strcpy(a.name, "C_crd"); a.dtype = 0; a.count = pos[n]; a.data = crd; outs.push_back(a);
        strcpy(a.name, "C_vals"); a.dtype = tbin_dtype<T>(); a.count = pos[n]; a.data = vals; outs.push_back(a);
        tbin_write(outPath, outs.data(), outs.size());
      }
    } else if (kernel == "pack_csr" || kernel == "pack_dcsr" || kernel == "pack_csf3" || kernel == "pack_csc") {
      // TensorBase::insert + pack() (src/tensor.cpp:295-463): COO entries (c0, c1[, c2], vals; any order, duplicates
      // allowed) -> level arrays of the target format.  assemble_ms = the inserts, compute_ms = pack().
      const int order = kernel == "pack_csf3" ? 3 : 2;
      std::vector<int> d(dims, dims + order);
      Format fmt = kernel == "pack_csr" ? Format({Dense, Sparse}) : kernel == "pack_csc" ? Format({Dense, Sparse}, {1, 0})
                   : kernel == "pack_dcsr" ? Format({Sparse, Sparse})
                                                                                           : Format({Sparse, Sparse, Sparse});
      Tensor<T> A("A", d, fmt);
      tbin_array* cv[3] = {need(in, "c0"), need(in, "c1"), order == 3 ? need(in, "c2") : nullptr};
      tbin_array* v = need(in, "vals");
      double t0 = now_ms();
      std::vector<int> c(order);
      for (uint64_t e = 0; e < v->count; e++) {
        for (int m = 0; m < order; m++) c[m] = ((int*)cv[m]->data)[e];
        A.insert(c, ((T*)v->data)[e]);
      }
      double t1 = now_ms();
      A.pack();
      double t2 = now_ms();
      if (rep == 0) tm.compile = 0;
      tm.assemble.push_back(t1 - t0); tm.compute.push_back(t2 - t1);
      if (last) {
        auto st = A.getStorage();
        auto idx = st.getIndex();
        std::vector<std::string> names;
        names.reserve(16);
        size_t parent = 1;
        for (int lv = 0; lv < order; lv++) {
          auto mi = idx.getModeIndex(lv);
          if (mi.numIndexArrays() < 2) { parent *= d[fmt.getModeOrdering()[lv]]; continue; }     // dense level
          Array pos = mi.getIndexArray(0), crd = mi.getIndexArray(1);
          int* pp = (int*)pos.getData();
          tbin_array a; a.dtype = 0;
          names.push_back("A" + std::to_string(lv + 1) + "_pos"); strcpy(a.name, names.back().c_str());
          a.count = parent + 1; a.data = pp; outs.push_back(a);
          names.push_back("A" + std::to_string(lv + 1) + "_crd"); strcpy(a.name, names.back().c_str());
          a.count = pp[parent]; a.data = crd.getData(); outs.push_back(a);
          parent = pp[parent];
        }
        tbin_array a; strcpy(a.name, "A_vals"); a.dtype = tbin_dtype<T>(); a.count = parent; a.data = st.getValues().getData();
        outs.push_back(a);
        tbin_write(outPath, outs.data(), outs.size());
      }
    } else if (kernel == "bspmv" || kernel == "bspmm") {
      // bspmv: a(i,j) = A(i,k,j,l) * c(k,l)   (the reference's blocked-SpMV test statement)
      // bspmm: C(i,j,m) = A(i,k,j,l) * B(k,l,m)
      int Mb = dims[0], Nb = dims[1], br = dims[2], bc = dims[3];
      IndexVar m("m");
      Tensor<T> A = attachBCSR<T>("A", {Mb, Nb, br, bc}, in, "A");
      if (kernel == "bspmv") {
        Tensor<T> c = attachDense<T>("c", {Nb, bc}, (T*)need(in, "c")->data);
        Tensor<T> a("a", {Mb, br}, Format({Dense, Dense}));
        a(i, j) = A(i, k, j, l) * c(k, l);
        double t0 = now_ms(); a.compile();
        double t1 = now_ms(); dumpSource(a); a.assemble();
        double t2 = now_ms(); a.compute();
        double t3 = now_ms();
        if (rep == 0) tm.compile = t1 - t0;
        tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
        if (last) {
          tbin_array o; strcpy(o.name, "a"); o.dtype = tbin_dtype<T>(); o.count = (uint64_t)Mb * br;
          o.data = a.getStorage().getValues().getData(); outs.push_back(o);
          tbin_write(outPath, outs.data(), outs.size());
        }
      } else {
        int K = dims[4];
        Tensor<T> B = attachDense<T>("B", {Nb, bc, K}, (T*)need(in, "B")->data);
        Tensor<T> C("C", {Mb, br, K}, Format({Dense, Dense, Dense}));
        C(i, j, m) = A(i, k, j, l) * B(k, l, m);
        IndexStmt stmt = C.getAssignment().concretize();
        if (tuned) {   // rows of blocks in parallel (the shape of scheduleSpMMCPU without the pos split)
          stmt = stmt.parallelize(i, ParallelUnit::CPUThread, OutputRaceStrategy::NoRaces);
        }
        double t0 = now_ms(); if (tuned) C.compile(stmt); else C.compile();
        double t1 = now_ms(); dumpSource(C); C.assemble();
        double t2 = now_ms(); C.compute();
        double t3 = now_ms();
        if (rep == 0) tm.compile = t1 - t0;
        tm.assemble.push_back(t2 - t1); tm.compute.push_back(t3 - t2);
        if (last) {
          tbin_array o; strcpy(o.name, "C"); o.dtype = tbin_dtype<T>(); o.count = (uint64_t)Mb * br * K;
          o.data = C.getStorage().getValues().getData(); outs.push_back(o);
          tbin_write(outPath, outs.data(), outs.size());
        }
      }
    } else {
      std::cerr << "unknown kernel " << kernel << std::endl;
      return 2;
    }
  }
  return 0;
}

// `taco_ref_harness emit_cuda <spmv|spmm> <outdir/> <prefix>`: print the CUDA source the REFERENCE generates for its own GPU
// schedule of the statement (scheduleSpMVGPU / scheduleSpMMGPU, test/tests-scheduling-eval.cpp:193-209, 249-268, same
// directives and parameters) through the reference's CodeGen_CUDA (Module::compileToSource).  No GPU or CUDA build is
// needed to GENERATE the text; oracle/Makefile compiles it with nvcc for sm_100a as the "recompiled reference kernel"
// comparator (tools/ref_cuda_bench.cu).  Output goes to oracle/_ref only.
template <typename T>
static int emit_cuda(const std::string& which, const std::string& dir, const std::string& prefix) {
  set_CUDA_codegen_enabled(true);
  IndexVar i("i"), j("j"), k("k");
  const int WARP = 32;
  IndexStmt stmt;
  if (which == "spmv") {
    Tensor<T> A("A", {1024, 1024}, CSR), x("x", {1024}, Format({Dense})), y("y", {1024}, Format({Dense}));
    IndexExpr pre = A(i, j) * x(j);
    y(i) = pre;
    stmt = y.getAssignment().concretize();
    const int NNZ_PER_THREAD = 8, BLOCK = 256;
    IndexVar f("f"), fpos("fpos"), fpos1("fpos1"), fpos2("fpos2"), block("block"), warp("warp"), thread("thread"),
        thread_nz("thread_nz"), thread_nz_pre("thread_nz_pre");
    TensorVar precomputed("precomputed", Type(type<T>(), {Dimension(thread_nz)}), taco::dense);
    stmt = stmt.fuse(i, j, f).pos(f, fpos, A(i, j)).split(fpos, block, fpos1, NNZ_PER_THREAD * BLOCK)
               .split(fpos1, warp, fpos2, NNZ_PER_THREAD * WARP).split(fpos2, thread, thread_nz, NNZ_PER_THREAD)
               .reorder({block, warp, thread, thread_nz}).precompute(pre, thread_nz, thread_nz_pre, precomputed)
               .unroll(thread_nz_pre, NNZ_PER_THREAD)
               .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
               .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
               .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
  } else if (which == "spmm") {
    Tensor<T> A("A", {1024, 1024}, CSR), B("B", {1024, 128}, Format({Dense, Dense})), C("C", {1024, 128}, Format({Dense, Dense}));
    C(i, k) = A(i, j) * B(j, k);
    stmt = C.getAssignment().concretize();
    const int NNZ_PER_WARP = 8, BLOCK = 256;
    IndexVar f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
        dense_val("dense_val"), thread("thread");
    stmt = stmt.reorder({i, j, k}).fuse(i, j, f).pos(f, fpos, A(i, j)).split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP))
               .split(fpos1, warp, nnz, NNZ_PER_WARP).split(k, dvu, thread, WARP).reorder({block, warp, thread, dvu, nnz})
               .bound(dvu, dense_val, 4, BoundType::MaxExact)
               .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
               .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
               .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
  } else if (which == "mttkrp") {           // scheduleMTTKRPGPU, tests-scheduling-eval.cpp:327-342 (rank 32 = one warp of columns)
    IndexVar l("l");
    Tensor<T> B("B", {64, 64, 64}, Format({Sparse, Sparse, Sparse})), C("C", {64, 32}, Format({Dense, Dense})),
        D("D", {64, 32}, Format({Dense, Dense})), A("A", {64, 32}, Format({Dense, Dense}));
    A(i, j) = B(i, k, l) * C(k, j) * D(l, j);
    stmt = A.getAssignment().concretize();
    const int NNZ_PER_WARP = 16, BLOCK = 256;
    IndexVar kl("kl"), f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
        dense_val("dense_val"), thread("thread");
    stmt = stmt.reorder({i, k, l, j}).fuse(k, l, kl).fuse(i, kl, f).pos(f, fpos, B(i, k, l))
               .split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP)).split(fpos1, warp, nnz, NNZ_PER_WARP)
               .split(j, dvu, thread, WARP).bound(dvu, dense_val, 1, BoundType::MaxExact).reorder({block, warp, dense_val, thread, nnz})
               .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
               .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
               .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
  } else if (which == "ttv") {              // scheduleTTVGPU, tests-scheduling-eval.cpp:308-325
    Tensor<T> B("B", {64, 64, 64}, Format({Sparse, Sparse, Sparse})), c("c", {64}, Format({Dense})), A("A", {64, 64}, Format({Dense, Dense}));
    IndexExpr pre = B(i, j, k) * c(k);
    A(i, j) = pre;
    stmt = A.getAssignment().concretize();
    const int NNZ_PER_WARP = 8 * 32, BLOCK = 256;
    IndexVar jk("jk"), f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), fpos2("fpos2"), thread("thread"),
        thread_nz("thread_nz"), thread_nz_pre("thread_nz_pre");
    TensorVar precomputed("precomputed", Type(type<T>(), {Dimension(thread_nz)}), taco::dense);
    stmt = stmt.fuse(j, k, jk).fuse(i, jk, f).pos(f, fpos, B(i, j, k)).split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP))
               .split(fpos1, warp, fpos2, NNZ_PER_WARP).split(fpos2, thread, thread_nz, NNZ_PER_WARP / WARP)
               .reorder({block, warp, thread, thread_nz}).precompute(pre, thread_nz, thread_nz_pre, precomputed)
               .unroll(thread_nz_pre, NNZ_PER_WARP / WARP)
               .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
               .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
               .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
  } else if (which == "ttm") {              // scheduleTTMGPU, tests-scheduling-eval.cpp:289-306, CO_FACTOR = 1 (32 result columns)
    IndexVar l("l");
    Tensor<T> B("B", {64, 64, 64}, Format({Sparse, Sparse, Sparse})), C("C", {64, 32}, Format({Dense, Dense})),
        A("A", {64, 64, 32}, Format({Dense, Dense, Dense}));
    A(i, j, l) = B(i, j, k) * C(k, l);
    stmt = A.getAssignment().concretize();
    const int NNZ_PER_WARP = 8 * 32, BLOCK = 256, CO_FACTOR = 1;
    IndexVar jk("jk"), f("f"), fpos("fpos"), block("block"), fpos1("fpos1"), warp("warp"), nnz("nnz"), dvu("dense_val_unbounded"),
        dense_val("dense_val"), thread("thread");
    stmt = stmt.reorder({i, j, k, l}).fuse(j, k, jk).fuse(i, jk, f).pos(f, fpos, B(i, j, k))
               .split(fpos, block, fpos1, NNZ_PER_WARP * (BLOCK / WARP)).split(fpos1, warp, nnz, NNZ_PER_WARP)
               .split(l, dvu, thread, WARP).bound(dvu, dense_val, CO_FACTOR, BoundType::MaxExact)
               .reorder({block, warp, nnz, thread, dense_val}).unroll(dense_val, CO_FACTOR)
               .parallelize(block, ParallelUnit::GPUBlock, OutputRaceStrategy::IgnoreRaces)
               .parallelize(warp, ParallelUnit::GPUWarp, OutputRaceStrategy::IgnoreRaces)
               .parallelize(thread, ParallelUnit::GPUThread, OutputRaceStrategy::Atomics);
  } else {
    std::cerr << "emit_cuda: spmv | spmm | mttkrp | ttv | ttm" << std::endl;
    return 2;
  }
  ir::Module module;
  module.addFunction(lower(stmt, "compute", false, true));
  module.compileToSource(dir, prefix);
  return 0;
}

int main(int argc, char** argv) {
  if (argc >= 5 && std::string(argv[1]) == "emit_cuda") {
    try {
      return std::string(argv[2]) == "spmm" ? emit_cuda<float>(argv[2], argv[3], argv[4]) : emit_cuda<double>(argv[2], argv[3], argv[4]);
    } catch (const TacoException& e) {
      std::cerr << "TacoException: " << e.what() << std::endl;
      return 3;
    }
  }
  if (argc < 4) {
    std::cerr << "usage: taco_ref_harness <kernel> <in.tbin> <out.tbin> [--dtype f64|f32] [--schedule default|cpu]"
                 " [--threads N] [--reps R]" << std::endl;
    return 2;
  }
  std::string kernel = argv[1], dtype = "f64", schedule = "default";
  int threads = 1, reps = 1;
  for (int a = 4; a + 1 < argc; a += 2) {
    std::string key = argv[a], val = argv[a + 1];
    if (key == "--dtype") dtype = val;
    else if (key == "--schedule") schedule = val;
    else if (key == "--threads") threads = atoi(val.c_str());
    else if (key == "--reps") reps = atoi(val.c_str());
    else if (key == "--no-out") g_no_out = atoi(val.c_str()) != 0;
  }
  tbin_file in;
  if (tbin_read(argv[2], &in) != 0) { std::cerr << "cannot read " << argv[2] << std::endl; return 2; }
  taco_set_num_threads(threads);
  Times tm;
  int rc = 1;
  try {
    rc = (dtype == "f32") ? run<float>(kernel, in, argv[3], schedule, reps, tm)
                          : run<double>(kernel, in, argv[3], schedule, reps, tm);
  } catch (const TacoException& e) {
    std::cerr << "TacoException: " << e.what() << std::endl;
    return 3;
  }
  std::cout << "{\"kernel\":\"" << kernel << "\",\"dtype\":\"" << dtype << "\",\"schedule\":\"" << schedule
            << "\",\"threads\":" << threads << ",\"compile_ms\":" << tm.compile << ",\"assemble_ms\":[";
  for (size_t r = 0; r < tm.assemble.size(); r++) std::cout << (r ? "," : "") << tm.assemble[r];
  std::cout << "],\"compute_ms\":[";
  for (size_t r = 0; r < tm.compute.size(); r++) std::cout << (r ? "," : "") << tm.compute[r];
  std::cout << "]}" << std::endl;
  return rc;
}
