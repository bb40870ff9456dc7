// examples/dropin_demo.cpp -- an ordinary taco C++ program (Format, Tensor<T>, IndexVar, assemble/compute) running its
// kernels on libtaco_b200 through the UNMODIFIED reference library.
//
// The only line that differs from a stock taco program is `T.compileSource(stub)` instead of `T.compile()`:
// compileSource (reference: src/tensor.cpp:905-930) is taco's documented plug point for user-supplied kernel text; the
// text here is the forwarding stub libtaco_b200 generates for the statement (taco_b200_module_stub_source), which taco
// JIT-compiles with its usual `cc` and calls through its usual _shim_ entry points (src/codegen/module.cpp:178-218).
// Every result is compared with taco's own C code generator (`compile()`) on the same operands in the same process.
//
//   build (oracle/Makefile, target ref):  g++ examples/dropin_demo.cpp -I$TACO_REF/include -Iinclude -Loracle/_ref -ltaco
//   run:  oracle/_ref/taco_dropin_demo taco_b200/lib/libtaco_b200.so
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>
#include <random>
#include <string>

#include "taco.h"

using namespace taco;

typedef struct taco_b200_module taco_b200_module_t;
static taco_b200_module_t* (*b200_open)(const char*, const char*, const char*);
static const char* (*b200_stub)(taco_b200_module_t*);
static const char* (*b200_err)(void);

static std::string stub_for(const char* expr, const char* formats, const char* dtype) {
  taco_b200_module_t* m = b200_open(expr, formats, dtype);
  if (!m) { fprintf(stderr, "module_open failed: %s\n", b200_err()); exit(2); }
  return b200_stub(m);
}

template <typename T>
static bool same(const Tensor<T>& a, const Tensor<T>& b, double rtol, const char* what) {
  // structure (pos/crd of every compressed level) must be identical, values within rtol
  bool ok = equals(a, b);
  if (!ok) {
    // equals() is exact on values; fall back to a tolerance walk over the stored components
    auto ia = a.begin(), ib = b.begin();
    ok = true;
    for (; ia != a.end() && ib != b.end(); ++ia, ++ib) {
      if (ia->first != ib->first) { ok = false; break; }
      double x = ia->second, y = ib->second;
      if (std::fabs(x - y) > rtol * std::fmax(std::fabs(x), std::fabs(y))) { ok = false; break; }
    }
    ok = ok && ia == a.end() && ib == b.end();
  }
  printf("%-8s %s\n", what, ok ? "PASS (matches taco's C codegen)" : "FAIL");
  return ok;
}

int main(int argc, char** argv) {
  const char* libpath = argc > 1 ? argv[1] : "libtaco_b200.so";
  setenv("TACO_B200_LIB", libpath, 1);
  setenv("TACO_CFLAGS", "-O3 -std=gnu99", 0);          // gnu99: the stub uses dlopen(); no -ffast-math for parity
  void* lib = dlopen(libpath, RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  b200_open = (taco_b200_module_t * (*)(const char*, const char*, const char*)) dlsym(lib, "taco_b200_module_open");
  b200_stub = (const char* (*)(taco_b200_module_t*))dlsym(lib, "taco_b200_module_stub_source");
  b200_err = (const char* (*)(void))dlsym(lib, "taco_b200_last_error");

  const int n = 2000, m = 1500, K = 24;
  std::mt19937 gen(7);
  std::uniform_real_distribution<double> val(0.0, 1.0);
  Format csr({Dense, Sparse}), dv({Dense}), dm({Dense, Dense});
  Tensor<double> A("A", {n, m}, csr), B("B", {n, m}, csr), S("S", {m, n}, csr), x("x", {m}, dv), D("D", {m, K}, dm);
  for (int i = 0; i < n; i++)
    for (int t = 0; t < 9; t++) {
      A.insert({i, (int)(gen() % m)}, std::floor(val(gen) * 1024) / 1024);
      if (t < 6) B.insert({i, (int)(gen() % m)}, std::floor(val(gen) * 1024) / 1024);
    }
  for (int i = 0; i < m; i++)
    for (int t = 0; t < 5; t++) S.insert({i, (int)(gen() % n)}, std::floor(val(gen) * 1024) / 1024);
  for (int j = 0; j < m; j++) {
    x.insert({j}, std::floor(val(gen) * 1024) / 1024);
    for (int k = 0; k < K; k++) D.insert({j, k}, std::floor(val(gen) * 1024) / 1024);
  }
  A.pack(); B.pack(); S.pack(); x.pack(); D.pack();
  IndexVar i("i"), j("j"), k("k");
  bool ok = true;

  {  // SpMV  y(i) = A(i,j) * x(j)
    Tensor<double> y("y", {n}, dv), yr("yr", {n}, dv);
    y(i) = A(i, j) * x(j);
    y.compileSource(stub_for("y(i) = A(i,j) * x(j)", "y:d,A:ds,x:d", "f64"));
    y.assemble(); y.compute();
    yr(i) = A(i, j) * x(j);
    yr.compile(); yr.assemble(); yr.compute();
    ok &= same(y, yr, 1e-12, "SpMV");
  }
  {  // SpMM  C(i,k) = A(i,j) * D(j,k)
    Tensor<double> C("C", {n, K}, dm), Cr("Cr", {n, K}, dm);
    C(i, k) = A(i, j) * D(j, k);
    C.compileSource(stub_for("C(i,k) = A(i,j) * D(j,k)", "C:dd,A:ds,D:dd", "f64"));
    C.assemble(); C.compute();
    Cr(i, k) = A(i, j) * D(j, k);
    Cr.compile(); Cr.assemble(); Cr.compute();
    ok &= same(C, Cr, 1e-12, "SpMM");
  }
  {  // SpAdd  C(i,j) = A(i,j) + B(i,j)   -- sparse result: taco reads pos[n] back and adopts our pos/crd/vals
    Tensor<double> C("C", {n, m}, csr), Cr("Cr", {n, m}, csr);
    C(i, j) = A(i, j) + B(i, j);
    C.compileSource(stub_for("C(i,j) = A(i,j) + B(i,j)", "C:ds,A:ds,B:ds", "f64"));
    C.assemble(); C.compute();
    Cr(i, j) = A(i, j) + B(i, j);
    Cr.compile(); Cr.assemble(); Cr.compute();
    ok &= same(C, Cr, 1e-12, "SpAdd");
  }
  {  // SpGEMM  C(i,k) = A(i,j) * S(j,k)
    Tensor<double> C("C", {n, n}, csr), Cr("Cr", {n, n}, csr);
    C(i, k) = A(i, j) * S(j, k);
    C.compileSource(stub_for("C(i,k) = A(i,j) * S(j,k)", "C:ds,A:ds,S:ds", "f64"));
    C.assemble(); C.compute();
    Cr(i, k) = A(i, j) * S(j, k);
    Cr.compile(); Cr.assemble(); Cr.compute();
    ok &= same(C, Cr, 1e-12, "SpGEMM");
  }
  printf("%s\n", ok ? "dropin_demo: ALL PASS" : "dropin_demo: FAILURES");
  return ok ? 0 : 1;
}
