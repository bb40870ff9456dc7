// module.cu -- the GPU-path replacement for taco's JIT module, plus the _shim_ entry points and the partitioner.
//
// Reference flow being replaced (/root/reference/src/codegen/module.cpp:111-218, src/tensor.cpp:605-674):
//   lower(stmt) -> CodeGen_CUDA prints a .cu -> system("nvcc ...") (5.8 s per statement) ->
//   dlopen -> dlsym("_shim_compute") -> callFuncPacked(void** args).
// Here there is no text generation and no compiler in the loop: a concrete index statement is CLASSIFIED
// (expression shape up to index-variable / tensor renaming, per-tensor level formats, component type) into one of
// the hand-written sm_100a kernel families, which are already resident in this library.  taco_b200_module_open() is
// the cache lookup (keyed on the canonical statement string), taco_b200_module_call_packed() is callFuncPacked().
// A statement that is not on the hot path is refused (TACO_B200_ERR_UNSUPPORTED) -- never run on the CPU.
#include <algorithm>
#include <cctype>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace tb {

typedef int (*fn3_t)(taco_tensor_t*, taco_tensor_t*, taco_tensor_t*);
typedef int (*fn4_t)(taco_tensor_t*, taco_tensor_t*, taco_tensor_t*, taco_tensor_t*);

struct Family {
  const char* name;
  const char* canon;                  // canonical expression: tensors T0.. (T0 = result), index vars a,b,c,d
  const char* formats[4];             // required level formats per tensor, in argument order ("d","ds","sss",...)
  int nargs;
  void* assemble; void* compute; void* evaluate;
  const char* arg1_ordering = nullptr;   // mode ordering the sparse operand (argument 1) must be stored in; null = identity
};

#define TB_FAM3(n) (void*)taco_b200_##n##_assemble, (void*)taco_b200_##n##_compute, (void*)taco_b200_##n##_evaluate

static const Family kFamilies[] = {
    {"spmv",   "T0(a)=T1(a,b)*T2(b)",                 {"d", "ds", "d", ""},      3, TB_FAM3(spmv)},
    {"spmm",   "T0(a,b)=T1(a,c)*T2(c,b)",             {"dd", "ds", "dd", ""},    3, TB_FAM3(spmm)},
    {"spmm_dcsr", "T0(a,b)=T1(a,c)*T2(c,b)",          {"dd", "ss", "dd", ""},    3, TB_FAM3(spmm_dcsr)},
    {"spgemm", "T0(a,b)=T1(a,c)*T2(c,b)",             {"ds", "ds", "ds", ""},    3, TB_FAM3(spgemm)},
    {"spadd",  "T0(a,b)=T1(a,b)+T2(a,b)",             {"ds", "ds", "ds", ""},    3, TB_FAM3(spadd)},
    {"sddmm",  "T0(a,b)=T1(a,b)*T2(a,c)*T3(b,c)",     {"ds", "ds", "dd", "dd"},  4, TB_FAM3(sddmm)},
    {"sddmm_dense", "T0(a,b)=T1(a,b)*T2(a,c)*T3(c,b)", {"dd", "ds", "dd", "dd"},  4, TB_FAM3(sddmm_dense)},
    {"mttkrp", "T0(a,b)=T1(a,c,d)*T2(c,b)*T3(d,b)",   {"dd", "sss", "dd", "dd"}, 4, TB_FAM3(mttkrp)},
    // mode-J / mode-K MTTKRP of a CP-ALS sweep (the reference's parafac tests, test/tests-parafac.cpp:157-187, factories
    // test/expr_factory.cpp:100-124): with B stored in the mode ordering that puts the result's mode first, its level arrays
    // are the CSF of the permuted tensor and the statement is the standard MTTKRP over the storage order
    {"mttkrp", "T0(a,b)=T1(c,a,d)*T2(c,b)*T3(d,b)",   {"dd", "sss", "dd", "dd"}, 4, TB_FAM3(mttkrp), "1,0,2"},
    {"mttkrp", "T0(a,b)=T1(c,d,a)*T2(c,b)*T3(d,b)",   {"dd", "sss", "dd", "dd"}, 4, TB_FAM3(mttkrp), "2,0,1"},
    {"ttv",    "T0(a,b)=T1(a,b,c)*T2(c)",             {"dd", "sss", "d", ""},    3, TB_FAM3(ttv)},
    {"ttm",    "T0(a,b,c)=T1(a,b,d)*T2(d,c)",         {"ddd", "sss", "dd", ""},  3, TB_FAM3(ttm)},
    {"bspmv",  "T0(a,b)=T1(a,c,b,d)*T2(c,d)",         {"dd", "dsdd", "dd", ""},  3, TB_FAM3(bspmv)},
    {"bspmm",  "T0(a,b,c)=T1(a,d,b,e)*T2(d,e,c)",     {"ddd", "dsdd", "ddd", ""}, 3, TB_FAM3(bspmm)},
};

// "y(i) = x(j) * A(i,j)"  ->  result access, operand accesses, operator
struct Acc { std::string name; std::vector<std::string> vars; };
struct Parsed { Acc lhs; std::vector<Acc> terms; char op = 0; };

static bool parse_expr(const std::string& expr, Parsed* out) {
  size_t p = 0;
  auto skip = [&]() { while (p < expr.size() && isspace((unsigned char)expr[p])) p++; };
  auto ident = [&](std::string* s) {
    skip();
    size_t b = p;
    while (p < expr.size() && (isalnum((unsigned char)expr[p]) || expr[p] == '_')) p++;
    *s = expr.substr(b, p - b);
    return p > b;
  };
  bool first = true;
  while (true) {
    Acc a;
    if (!ident(&a.name)) return false;
    skip();
    if (p < expr.size() && expr[p] == '(') {
      p++;
      while (true) {
        std::string v;
        if (!ident(&v)) return false;
        a.vars.push_back(v);
        skip();
        if (p < expr.size() && expr[p] == ',') { p++; continue; }
        if (p < expr.size() && expr[p] == ')') { p++; break; }
        return false;
      }
    }
    if (first) out->lhs = a; else out->terms.push_back(a);
    skip();
    if (p >= expr.size()) break;
    const char c = expr[p];
    if (first && c == '=') { p++; first = false; continue; }
    if (first && c == '+' && p + 1 < expr.size() && expr[p + 1] == '=') { p += 2; first = false; continue; }
    if (!first && (c == '*' || c == '+')) {
      if (out->op && out->op != c) return false;          // products or sums, not mixtures
      out->op = c;
      p++;
      continue;
    }
    return false;
  }
  return !first && !out->terms.empty();
}

// canonical text of the statement with its operands taken in the order `order`: tensors T0.. (T0 = result) and index
// variables a,b,c.. are numbered by first appearance; `names` receives the tensor names in that numbering
static std::string canonical(const Parsed& ps, const std::vector<int>& order, std::vector<std::string>* names) {
  std::map<std::string, int> tid, vid;
  names->clear();
  std::string out;
  auto emit = [&](const Acc& a) {
    if (!tid.count(a.name)) { const int id = (int)tid.size(); tid[a.name] = id; names->push_back(a.name); }
    out += "T" + std::to_string(tid[a.name]);
    if (!a.vars.empty()) {
      out += "(";
      for (size_t v = 0; v < a.vars.size(); v++) {
        if (!vid.count(a.vars[v])) { const int id = (int)vid.size(); vid[a.vars[v]] = id; }
        out += (v ? "," : "");
        out += (char)('a' + vid[a.vars[v]]);
      }
      out += ")";
    }
  };
  emit(ps.lhs);
  out += "=";
  for (size_t t = 0; t < order.size(); t++) {
    if (t) out += ps.op ? ps.op : '*';
    emit(ps.terms[order[t]]);
  }
  return out;
}

// "A:ds,x:d,C:dd:1,0" -> per tensor (levels, ordering string)
static void parse_formats(const char* formats, std::map<std::string, std::pair<std::string, std::string>>* out) {
  if (!formats) return;
  std::string s(formats);
  size_t p = 0;
  while (p < s.size()) {
    size_t c1 = s.find(':', p);
    if (c1 == std::string::npos) break;
    std::string name = s.substr(p, c1 - p);
    size_t q = c1 + 1;
    std::string lv;
    while (q < s.size() && isalpha((unsigned char)s[q])) lv += s[q++];
    std::string ord;
    if (q < s.size() && s[q] == ':') {
      q++;
      while (q < s.size() && (isdigit((unsigned char)s[q]) || (s[q] == ',' && q + 1 < s.size() && isdigit((unsigned char)s[q + 1]))))
        ord += s[q++];
    }
    while (!name.empty() && isspace((unsigned char)name[0])) name.erase(0, 1);
    (*out)[name] = {lv, ord};
    p = q;
    if (p < s.size() && s[p] == ',') p++;
  }
}

}  // namespace tb

struct taco_b200_module {
  const tb::Family* fam;
  std::string key;
  int perm[4];           // family argument a is the caller's argument perm[a]
  bool identity;
  std::string stub;      // lazily built C source for TensorBase::compileSource()
};

using namespace tb;

static std::mutex g_mod_mu;
static std::map<std::string, taco_b200_module*> g_mod_cache;   // canonical statement -> module (process lifetime)

extern "C" {

// The operands of a product (or sum) commute: `y(i) = x(j) * A(i,j)` is SpMV with its arguments in another order.  Every
// order of the operands is tried against the family table; the module remembers how the caller's argument pack -- results
// first, then operands by first appearance, as taco packs them (src/tensor.cpp:778-806), or the explicit `args` list -- maps
// onto the family's signature.  (Reordering a product changes the association of floating-point products: permuted SDDMM /
// MTTKRP statements are within the tolerance of the reference's order, not bit-identical to it.)
taco_b200_module_t* taco_b200_module_open_args(const char* expr, const char* formats, const char* dtype, const char* args) {
  if (!expr) { fail(TACO_B200_ERR_ARG, "module_open: NULL expression"); return nullptr; }
  std::string dt = dtype ? dtype : "f64";
  if (dt == "float") dt = "f32";
  if (dt == "double") dt = "f64";
  if (dt != "f32" && dt != "f64") {
    fail(TACO_B200_ERR_UNSUPPORTED, "component type '%s' is not on the GPU hot path (f32 / f64 only)", dt.c_str());
    return nullptr;
  }
  Parsed ps;
  if (!parse_expr(expr, &ps) || ps.terms.size() > 3) {
    fail(TACO_B200_ERR_UNSUPPORTED, "cannot parse '%s' as  result(vars) = access {*|+} access ...", expr);
    return nullptr;
  }
  std::vector<int> order(ps.terms.size());
  for (size_t t = 0; t < order.size(); t++) order[t] = (int)t;
  std::vector<std::string> given;                       // the caller's argument order
  const std::string canon_given = canonical(ps, order, &given);
  if (args && *args) {
    given.clear();
    std::string a(args), cur;
    for (size_t q = 0; q <= a.size(); q++) {
      if (q == a.size() || a[q] == ',') {
        while (!cur.empty() && isspace((unsigned char)cur.back())) cur.pop_back();
        while (!cur.empty() && isspace((unsigned char)cur[0])) cur.erase(0, 1);
        if (!cur.empty()) given.push_back(cur);
        cur.clear();
      } else cur += a[q];
    }
  }
  std::map<std::string, std::pair<std::string, std::string>> fm;
  parse_formats(formats, &fm);
  std::string key = canon_given + "|" + (formats ? formats : "") + "|" + dt + "|";
  for (const std::string& g : given) key += g + ",";
  {
    std::lock_guard<std::mutex> lk(g_mod_mu);
    auto it = g_mod_cache.find(key);
    if (it != g_mod_cache.end()) return it->second;
  }
  std::string tried;
  do {
    std::vector<std::string> tensors;
    const std::string canon = canonical(ps, order, &tensors);
    if (tried.find(canon) == std::string::npos) tried += (tried.empty() ? "" : " | ") + canon;
    for (const Family& f : kFamilies) {
      if (canon != f.canon || (int)tensors.size() != f.nargs) continue;
      bool ok = true;
      for (int a = 0; a < f.nargs && ok; a++) {
        auto it = fm.find(tensors[a]);
        std::string lv = it == fm.end() ? std::string(strlen(f.formats[a]), 'd') : it->second.first;   // unlisted tensors are dense
        if (lv != f.formats[a]) ok = false;
        if (ok && a == 1 && f.arg1_ordering) {       // this family needs the sparse operand in a specific (non-identity) ordering
          if (it == fm.end() || it->second.second != f.arg1_ordering) ok = false;
          continue;
        }
        if (ok && it != fm.end() && !it->second.second.empty()) {
          // only the spmm result may carry a non-identity mode ordering (the reference GPU test's column-major C)
          std::string ident;
          for (size_t l = 0; l < lv.size(); l++) ident += (l ? "," : "") + std::to_string(l);
          if (it->second.second != ident && !((std::string(f.name) == "spmm" || std::string(f.name) == "spmm_dcsr") && a == 0 && it->second.second == "1,0")) ok = false;
        }
      }
      if (!ok) continue;
      taco_b200_module* m = new taco_b200_module{&f, key, {0, 1, 2, 3}, true, ""};
      bool mapped = (int)given.size() == f.nargs;
      for (int a = 0; a < f.nargs && mapped; a++) {
        int at = -1;
        for (int g = 0; g < f.nargs; g++) if (given[g] == tensors[a]) at = g;
        if (at < 0) mapped = false;
        else { m->perm[a] = at; m->identity &= (at == a); }
      }
      if (!mapped) {
        delete m;
        fail(TACO_B200_ERR_ARG, "module_open: the argument list '%s' does not name the %d tensors of '%s'", args ? args : "", f.nargs, expr);
        return nullptr;
      }
      std::lock_guard<std::mutex> lk(g_mod_mu);
      g_mod_cache[key] = m;
      return m;
    }
  } while (std::next_permutation(order.begin(), order.end()));
  fail(TACO_B200_ERR_UNSUPPORTED,
       "statement '%s' with formats '%s' is not a GPU hot-path pattern (canonical forms tried: %s); no CPU fallback exists",
       expr, formats ? formats : "", tried.c_str());
  return nullptr;
}

taco_b200_module_t* taco_b200_module_open(const char* expr, const char* formats, const char* dtype) {
  return taco_b200_module_open_args(expr, formats, dtype, nullptr);
}

const char* taco_b200_module_family(const taco_b200_module_t* m) { return m ? m->fam->name : nullptr; }
int taco_b200_module_num_args(const taco_b200_module_t* m) { return m ? m->fam->nargs : 0; }

static void* module_func(taco_b200_module_t* m, const char* name) {
  if (!m || !name) return nullptr;
  if (!strcmp(name, "assemble")) return m->fam->assemble;
  if (!strcmp(name, "compute")) return m->fam->compute;
  if (!strcmp(name, "evaluate")) return m->fam->evaluate;
  return nullptr;
}

void* taco_b200_module_get_func_ptr(taco_b200_module_t* m, const char* name) {
  if (m && !m->identity) {        // a raw entry point cannot reorder its arguments: use call_packed or the stub source
    fail(TACO_B200_ERR_UNSUPPORTED, "module_get_func_ptr: the statement's operands are permuted with respect to the %s kernel", m->fam->name);
    return nullptr;
  }
  return module_func(m, name);
}

int taco_b200_module_call_packed(taco_b200_module_t* m, const char* name, void** args) {
  if (!m) return fail(TACO_B200_ERR_ARG, "module_call_packed: NULL module");
  void* f = module_func(m, name);
  if (!f) return fail(TACO_B200_ERR_ARG, "module has no function '%s'", name ? name : "(null)");
  if (!args) return fail(TACO_B200_ERR_ARG, "module_call_packed: NULL argument pack");
  const int* p = m->perm;
  if (m->fam->nargs == 3)
    return ((fn3_t)f)((taco_tensor_t*)args[p[0]], (taco_tensor_t*)args[p[1]], (taco_tensor_t*)args[p[2]]);
  return ((fn4_t)f)((taco_tensor_t*)args[p[0]], (taco_tensor_t*)args[p[1]], (taco_tensor_t*)args[p[2]], (taco_tensor_t*)args[p[3]]);
}

void taco_b200_module_close(taco_b200_module_t*) { /* modules are cached for the process lifetime */ }

// C source that the UNMODIFIED reference can take through its own plug point TensorBase::compileSource(std::string)
// (/root/reference/src/tensor.cpp:905-930; CLI: -read-source, tools/taco.cpp:1212-1259): it defines `assemble`,
// `compute` and `evaluate` with the signatures taco's generated shims call (CodeGen_C::generateShim,
// src/codegen/codegen_c.cpp:591-628) and forwards them to this library, located through dlopen
// ($TACO_B200_LIB, else "libtaco_b200.so" on the loader path).  taco compiles it with its normal `cc` JIT -- no nvcc, no
// change to taco.
const char* taco_b200_module_stub_source(taco_b200_module_t* m) {
  if (!m) { fail(TACO_B200_ERR_ARG, "module_stub_source: NULL module"); return nullptr; }
  std::lock_guard<std::mutex> lk(g_mod_mu);
  if (!m->stub.empty()) return m->stub.c_str();
  const int n = m->fam->nargs;
  std::string params, args, types;
  for (int a = 0; a < n; a++) {
    params += std::string(a ? ", " : "") + "taco_tensor_t* t" + std::to_string(a);
    args += std::string(a ? ", " : "") + "t" + std::to_string(m->perm[a]);      // the caller's order -> the kernel's order
    types += std::string(a ? ", " : "") + "taco_tensor_t*";
  }
  std::string src =
      "// generated by libtaco_b200 (taco_b200_module_stub_source): forwards taco's kernel entry points to the GPU path\n"
      "#include <stdint.h>\n#include <stdio.h>\n#include <stdlib.h>\n#include <dlfcn.h>\n"
      "#ifndef TACO_TENSOR_T_DEFINED\n#define TACO_TENSOR_T_DEFINED\n"
      "typedef enum { taco_mode_dense, taco_mode_sparse } taco_mode_t;\n"
      "typedef struct taco_tensor_t {\n  int32_t order; int32_t* dimensions; int32_t csize; int32_t* mode_ordering;\n"
      "  taco_mode_t* mode_types; uint8_t*** indices; uint8_t* vals; uint8_t* fill_value; int32_t vals_size;\n"
      "} taco_tensor_t;\n#endif\n"
      "static void* taco_b200_sym(const char* name) {\n"
      "  static void* lib = 0;\n"
      "  if (!lib) {\n"
      "    const char* path = getenv(\"TACO_B200_LIB\");\n"
      "    lib = dlopen(path ? path : \"libtaco_b200.so\", RTLD_NOW | RTLD_GLOBAL);\n"
      "    if (!lib) { fprintf(stderr, \"taco_b200: %s\\n\", dlerror()); abort(); }\n"
      "  }\n"
      "  void* f = dlsym(lib, name);\n"
      "  if (!f) { fprintf(stderr, \"taco_b200: missing symbol %s\\n\", name); abort(); }\n"
      "  return f;\n}\n"
      "static int taco_b200_report(int rc) {\n"
      "  if (rc) fprintf(stderr, \"taco_b200: %s\\n\", ((const char* (*)(void))taco_b200_sym(\"taco_b200_last_error\"))());\n"
      "  return rc;\n}\n";
  for (const char* ph : {"assemble", "compute", "evaluate"}) {
    src += std::string("int ") + ph + "(" + params + ") {\n  typedef int (*fn_t)(" + types + ");\n  static fn_t fn = 0;\n" +
           "  if (!fn) fn = (fn_t)taco_b200_sym(\"taco_b200_" + m->fam->name + "_" + ph + "\");\n" +
           "  return taco_b200_report(fn(" + args + "));\n}\n";
  }
  m->stub = src;
  return m->stub.c_str();
}

// ---- _shim_ entry points (positional void** pack, /root/reference/src/codegen/codegen_cuda.cpp:1500-1540) -------
#define TB_SHIM3(n, ph)                                                                                     \
  int _shim_taco_b200_##n##_##ph(void** p) {                                                                \
    return taco_b200_##n##_##ph((taco_tensor_t*)p[0], (taco_tensor_t*)p[1], (taco_tensor_t*)p[2]);          \
  }
#define TB_SHIM4(n, ph)                                                                                     \
  int _shim_taco_b200_##n##_##ph(void** p) {                                                                \
    return taco_b200_##n##_##ph((taco_tensor_t*)p[0], (taco_tensor_t*)p[1], (taco_tensor_t*)p[2], (taco_tensor_t*)p[3]); \
  }
#define TB_SHIMS3(n) TB_SHIM3(n, assemble) TB_SHIM3(n, compute) TB_SHIM3(n, evaluate)
#define TB_SHIMS4(n) TB_SHIM4(n, assemble) TB_SHIM4(n, compute) TB_SHIM4(n, evaluate)
TB_SHIMS3(spmv) TB_SHIMS3(spmm) TB_SHIMS3(spmm_dcsr) TB_SHIMS4(sddmm) TB_SHIMS4(sddmm_dense) TB_SHIMS4(mttkrp) TB_SHIMS3(ttv) TB_SHIMS3(ttm) TB_SHIMS3(spadd) TB_SHIMS3(spgemm) TB_SHIMS3(bspmv) TB_SHIMS3(bspmm)

}  // extern "C"

// ---- partitioner ------------------------------------------------------------------------------------------
namespace tb {
__global__ void partition_kernel(const int* __restrict__ pos, int parent, int parts, int* __restrict__ bounds) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g > parts) return;
  if (g == 0) { bounds[0] = 0; return; }
  if (g == parts) { bounds[parts] = parent; return; }
  long long total = pos[parent];
  int target = (int)((total * g) / parts);
  bounds[g] = tbd::search_first_ge(pos, 0, parent, target);
}
}  // namespace tb

extern "C" int taco_b200_partition_pos(const int32_t* pos, int32_t parent_size, int32_t parts, int32_t* bounds) {
  if (!pos || !bounds || parts <= 0 || parent_size < 0) return fail(TACO_B200_ERR_ARG, "partition_pos: bad argument");
  if (classify(pos) == Mem::Device) {
    TB_TRY(ensure_init());
    void* d = nullptr;
    TB_TRY(scratch_alloc(&d, sizeof(int) * (size_t)(parts + 1)));
    partition_kernel<<<(parts + 1 + 127) / 128, 128, 0, stream()>>>(pos, parent_size, parts, (int*)d);
    count_launch(1);
    int rc = read_back(bounds, d, sizeof(int) * (size_t)(parts + 1));
    scratch_free(d);
    return rc;
  }
  // host-described tensor: the same search on the host (no device needed to plan a distribution)
  long long total = pos[parent_size];
  bounds[0] = 0;
  bounds[parts] = parent_size;
  for (int g = 1; g < parts; g++) {
    int target = (int)((total * g) / parts);
    int lo = 0, end = parent_size + 1;
    while (lo < end) { int mid = lo + ((end - lo) >> 1); if (pos[mid] >= target) end = mid; else lo = mid + 1; }
    bounds[g] = lo > parent_size ? parent_size : lo;
  }
  return TACO_B200_OK;
}
