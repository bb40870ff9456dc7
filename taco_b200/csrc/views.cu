// views.cu -- validation / unpacking of taco_tensor_t into typed views.
//
// Mirrors what the reference's generated code does at the top of every kernel
// (CodeGen::unpackTensorProperty, /root/reference/src/codegen/codegen.cpp:232-272):
//   A1_dimension = A->dimensions[0];  A2_pos = (int*)A->indices[1][0];  A2_crd = (int*)A->indices[1][1]; ...
// but checks the format first: the generated code is specialised per statement, this library is not.
#include "common.cuh"

namespace tb {

int dtype_of(const taco_tensor_t* t, DType* out) {
  if (t->csize == 32) { *out = DType::F32; return TACO_B200_OK; }
  if (t->csize == 64) { *out = DType::F64; return TACO_B200_OK; }
  return fail(TACO_B200_ERR_FORMAT, "component size %d bits is not supported (float32 / float64 only)", t->csize);
}

int view_dense(const taco_tensor_t* t, int order, const char* name, DenseView* v) {
  if (!t) return fail(TACO_B200_ERR_ARG, "%s: NULL tensor", name);
  if (t->order != order) return fail(TACO_B200_ERR_FORMAT, "%s: order %d, expected %d", name, t->order, order);
  if (order > 3) return fail(TACO_B200_ERR_FORMAT, "%s: dense order > 3 not supported", name);
  v->order = order;
  for (int l = 0; l < order; l++) {
    if (t->mode_types[l] != taco_mode_dense) return fail(TACO_B200_ERR_FORMAT, "%s: level %d is not dense", name, l);
    v->mode_order[l] = t->mode_ordering[l];
    v->dim[l] = t->dimensions[l];
    if (v->dim[l] < 0) return fail(TACO_B200_ERR_ARG, "%s: negative dimension", name);
  }
  v->vals = t->vals;
  return dtype_of(t, &v->dt);
}

int view_csr(const taco_tensor_t* t, const char* name, CsrView* v) {
  if (!t) return fail(TACO_B200_ERR_ARG, "%s: NULL tensor", name);
  if (t->order != 2 || t->mode_types[0] != taco_mode_dense || t->mode_types[1] != taco_mode_sparse ||
      t->mode_ordering[0] != 0 || t->mode_ordering[1] != 1)
    return fail(TACO_B200_ERR_FORMAT, "%s: expected CSR ({Dense,Compressed}, mode ordering 0,1)", name);
  v->rows = t->dimensions[0];
  v->cols = t->dimensions[1];
  if (v->rows < 0 || v->cols < 0) return fail(TACO_B200_ERR_ARG, "%s: negative dimension", name);
  v->pos = t->indices && t->indices[1] ? (int32_t*)t->indices[1][0] : nullptr;
  v->crd = t->indices && t->indices[1] ? (int32_t*)t->indices[1][1] : nullptr;
  v->vals = t->vals;
  return dtype_of(t, &v->dt);
}

int view_csf3(const taco_tensor_t* t, const char* name, Csf3View* v) {
  if (!t) return fail(TACO_B200_ERR_ARG, "%s: NULL tensor", name);
  if (t->order != 3) return fail(TACO_B200_ERR_FORMAT, "%s: expected an order-3 CSF tensor", name);
  bool seen[3] = {false, false, false};
  for (int l = 0; l < 3; l++) {
    // any mode ordering: level l stores mode mode_ordering[l]; the kernels work in STORAGE order (the module classifier
    // only pairs a permuted ordering with the statement whose result mode comes first in storage, module.cu)
    const int m = t->mode_ordering[l];
    if (t->mode_types[l] != taco_mode_sparse || m < 0 || m > 2 || seen[m])
      return fail(TACO_B200_ERR_FORMAT, "%s: expected CSF ({Compressed x3}, a permutation as mode ordering)", name);
    seen[m] = true;
    v->dim[l] = t->dimensions[m];
    v->pos[l] = (int32_t*)t->indices[l][0];
    v->crd[l] = (int32_t*)t->indices[l][1];
    if (!v->pos[l]) return fail(TACO_B200_ERR_ARG, "%s: level %d has no pos array", name, l);
  }
  v->vals = t->vals;
  return dtype_of(t, &v->dt);
}

}  // namespace tb
