"""pytaco-style front end over the GPU hot path, with ZERO-COPY device tensors.

Mirror of the part of the reference's Python bindings a hot-path user touches
(/root/reference/python_bindings/pytaco/pytensor/taco_tensor.py: from_array :583, from_sp_csr :527, from_sp_csc :555,
to_array :659, to_sp_csr :717, evaluate :2825, matmul :2378; C++ side python_bindings/src/pyTensor.cpp:49-135).

What the reference does under CUDA: `fromNpArr` copies every dense operand into `cudaMallocManaged` memory
(pyTensor.cpp:58-63) and `fromSpMatrix` is `taco_not_supported_yet` (:123-124) -- sparse operands cannot be handed to the
GPU at all.  Here any array that already lives in HBM is ATTACHED, not copied:
  * torch CUDA tensors, and any object exposing `__cuda_array_interface__` (CuPy, Numba, RAPIDS) or `__dlpack__`,
    become operands whose taco_tensor_t points at the producer's memory;
  * results computed in device space come back as objects that export `__cuda_array_interface__` and `__dlpack__`,
    so torch / CuPy consume them without a copy either;
  * host arrays (numpy, scipy.sparse) still work -- the library stages them, as for every host-described tensor.
"""
import numpy as np

from . import tensor as _t
from ._lib import TacoError

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

dense, compressed = _t.dense, _t.compressed
csr = _t.CSR
csc = _t.Format([_t.dense, _t.compressed], [1, 0])


def _as_device_or_host(a):
    """numpy array -> itself; torch tensor -> itself; CUDA array interface / DLPack producer -> zero-copy torch view"""
    if isinstance(a, np.ndarray) or (torch is not None and isinstance(a, torch.Tensor)):
        return a
    if torch is None:
        raise TacoError(2, "device arrays need torch for the zero-copy view")
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device="cuda")            # shares the producer's memory
    if hasattr(a, "__dlpack__"):
        return torch.from_dlpack(a)
    return np.asarray(a)


def _is_dev(a):
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


def _check_dtype(a):
    dt = np.dtype(str(a.dtype).replace("torch.", ""))
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TacoError(2, f"component type {dt} is not on the GPU hot path (float32 / float64)")
    return dt


def _index32(a, copy_ok=True):
    """int32 index array in the same space (pytaco accepts int32 / int64 indices; the ABI is int32)"""
    if _is_dev(a):
        return a if a.dtype == torch.int32 else a.to(torch.int32)
    a = np.asarray(a)
    return np.ascontiguousarray(a, dtype=np.int32)


class tensor:
    """A pytaco-like tensor: a taco_b200.Tensor plus the array objects that own its memory."""

    def __init__(self, t):
        self._t = t

    # --- pytaco properties ---------------------------------------------------------------------------------------
    @property
    def shape(self):
        return list(self._t.dims)

    @property
    def order(self):
        return len(self._t.dims)

    @property
    def dtype(self):
        return self._t.dtype

    @property
    def format(self):
        return self._t.format

    @property
    def name(self):
        return self._t.name

    def on_device(self):
        v = self._t.vals()
        return _is_dev(v)

    # --- export ----------------------------------------------------------------------------------------------------
    def values(self):
        """the stored values in their own space: numpy array (host) or torch CUDA tensor sharing the library's memory"""
        return self._t.vals()

    def level_arrays(self, level):
        return self._t.level(level)

    def __dlpack__(self, stream=None):
        v = self._dense_view()
        return v.__dlpack__() if stream is None else v.__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return self._dense_view().__dlpack_device__()

    @property
    def __cuda_array_interface__(self):
        v = self._dense_view()
        if not _is_dev(v):
            raise AttributeError("host tensor")
        return v.__cuda_array_interface__

    def _dense_view(self):
        if compressed in self._t.format.levels:
            raise TacoError(2, "only dense tensors export an array view (use to_sp_csr / level_arrays for sparse ones)")
        v = self._t.vals()
        shape = [self._t.dims[m] for m in self._t.format.ordering]       # storage order
        if _is_dev(v):
            v = v.view(*shape)
            return v.permute(*np.argsort(self._t.format.ordering).tolist()) if self._t.format.ordering != sorted(self._t.format.ordering) else v
        v = np.asarray(v).reshape(shape)
        return v.transpose(np.argsort(self._t.format.ordering)) if self._t.format.ordering != sorted(self._t.format.ordering) else v

    def to_array(self):
        """dense tensor -> numpy array (device results are copied to the host here, and only here)"""
        v = self._dense_view()
        return v.detach().cpu().numpy() if _is_dev(v) else np.array(v)

    toarray = to_array

    def to_torch(self):
        """dense tensor -> torch tensor sharing the tensor's memory (no copy for device tensors)"""
        v = self._dense_view()
        return v if _is_dev(v) else torch.from_numpy(np.ascontiguousarray(v))

    def to_sp_csr(self):
        """CSR tensor -> scipy.sparse.csr_matrix on the host (pytaco: to_sp_csr :717)"""
        import scipy.sparse as sp
        if self._t.format != _t.CSR:
            raise TacoError(2, "to_sp_csr needs a {dense, compressed} tensor")
        pos, crd = self._t.level(1)
        h = lambda a: a.detach().cpu().numpy() if _is_dev(a) else np.asarray(a)
        return sp.csr_matrix((h(self._t.vals()), h(crd), h(pos)), shape=tuple(self._t.dims))


def from_array(array, copy=False, name="t"):
    """dense operand from a numpy array, a torch tensor, or any `__cuda_array_interface__` / `__dlpack__` producer.
    Device arrays are attached zero-copy (copy=True clones them first, as pytaco's default does for host arrays)."""
    a = _as_device_or_host(array)
    dt = _check_dtype(a)
    if _is_dev(a) or (torch is not None and isinstance(a, torch.Tensor)):
        if not a.is_contiguous() or copy:
            a = a.contiguous().clone() if copy else a.contiguous()
        dims = list(a.shape)
        flat = a.view(-1)
        if not a.is_cuda:
            flat = flat.numpy()
    else:
        a = np.ascontiguousarray(a) if not copy else np.array(a, order="C")
        dims = list(a.shape)
        flat = a.reshape(-1)
    t = _t.Tensor(name, dims, _t.Format([dense] * len(dims)), dt)
    t.set_vals(flat)
    t._keep.append(a)
    return tensor(t)


def _from_compressed(indptr, indices, data, shape, ordering, name):
    indptr, indices, data = (_as_device_or_host(x) for x in (indptr, indices, data))
    dt = _check_dtype(data)
    pos, crd = _index32(indptr), _index32(indices)
    if _is_dev(data):
        data = data.contiguous()
    else:
        data = np.ascontiguousarray(data)
    t = _t.Tensor(name, list(shape), _t.Format([dense, compressed], ordering), dt)
    t.set_level(1, pos, crd)
    t.set_vals(data)
    t._keep += [pos, crd, data]
    return tensor(t)


def from_sp_csr(matrix, shape=None, name="t"):
    """CSR operand from a scipy.sparse.csr_matrix (host) or from an (indptr, indices, data) triple of device arrays
    (torch CUDA tensors / CuPy arrays / anything with `__cuda_array_interface__`): the triple is attached zero-copy when
    the indices are already int32.  The reference cannot do this under CUDA at all (pyTensor.cpp:123-124)."""
    if isinstance(matrix, (tuple, list)):
        indptr, indices, data = matrix
        if shape is None:
            raise TacoError(3, "from_sp_csr((indptr, indices, data)) needs shape=")
        return _from_compressed(indptr, indices, data, shape, None, name)
    return _from_compressed(matrix.indptr, matrix.indices, matrix.data, matrix.shape, None, name)


def from_sp_csc(matrix, shape=None, name="t"):
    """CSC operand ({dense, compressed} with mode ordering 1,0), same sources as from_sp_csr"""
    if isinstance(matrix, (tuple, list)):
        indptr, indices, data = matrix
        if shape is None:
            raise TacoError(3, "from_sp_csc((indptr, indices, data)) needs shape=")
        return _from_compressed(indptr, indices, data, shape, [1, 0], name)
    return _from_compressed(matrix.indptr, matrix.indices, matrix.data, matrix.shape, [1, 0], name)


def as_tensor(obj, name="t"):
    return obj if isinstance(obj, tensor) else from_array(obj, name=name)


def to_array(t):
    return t.to_array()


def to_sp_csr(t):
    return t.to_sp_csr()


def evaluate(expr, *operands, out_format=None, dtype=None, shape=None):
    """pytaco.evaluate (taco_tensor.py:2825): `expr` is index notation over the operand names in order of appearance, e.g.
    evaluate("C(i,k) = A(i,j) * B(j,k)", A, B).  The result lives where the operands live: device operands -> device result
    (no copies), host operands -> host result.  `out_format` defaults to all-dense; `shape` is inferred from the operands."""
    import re
    ops = [as_tensor(o) for o in operands]
    lhs, rhs = expr.split("=", 1)
    accesses = re.findall(r"([A-Za-z_]\w*)\s*\(([^)]*)\)", rhs)
    res_name, res_vars = re.match(r"\s*([A-Za-z_]\w*)\s*\(([^)]*)\)", lhs.replace("+", "")).groups()
    res_vars = [v.strip() for v in res_vars.split(",")]
    names = []
    for n, _ in accesses:
        if n not in names:
            names.append(n)
    if len(names) != len(ops):
        raise TacoError(3, f"'{expr}' names {len(names)} operands, {len(ops)} given")
    extent = {}
    for (n, vs) in accesses:
        t = ops[names.index(n)]
        for m, v in enumerate(x.strip() for x in vs.split(",")):
            extent[v] = t.shape[m]
    dims = shape if shape is not None else [extent[v] for v in res_vars]
    fmt = out_format if out_format is not None else _t.Format([dense] * len(dims))
    fmt = fmt if isinstance(fmt, _t.Format) else _t.Format(fmt)
    dt = np.dtype(dtype) if dtype is not None else ops[0].dtype
    res = _t.Tensor(res_name, dims, fmt, dt)
    for n, o in zip(names, ops):
        o._t.name = n
    on_dev = any(o.on_device() for o in ops)
    if on_dev and torch is not None:
        # device operands were (and the result will be) produced / consumed by the caller's torch stream: enqueue the library's
        # work on that same stream, so that neither side reads an array the other is still writing (found by running this
        # front end's test under compute-sanitizer, whose slowdown exposed the race with the library's private stream)
        _t.use_torch_stream()
    prev = _t.lib.taco_b200_get_result_space()
    _t.set_result_space("device" if on_dev else "host")
    try:
        k = _t.compile(expr, res, *[o._t for o in ops])
        k(res, *[o._t for o in ops])
    finally:
        _t.check(_t.lib.taco_b200_set_result_space(prev))
    out = tensor(res)
    out._operands = ops
    return out


def matmul(t1, t2, out_format=None, dtype=None):
    """pytaco.matmul (taco_tensor.py:2378) for the hot-path shapes: sparse x dense (SpMM), sparse x vector (SpMV),
    sparse x sparse with a sparse result (SpGEMM, out_format=csr)"""
    a, b = as_tensor(t1), as_tensor(t2)
    if b.order == 1:
        return evaluate("y(i) = A(i,j) * x(j)", a, b, out_format=out_format, dtype=dtype)
    return evaluate("C(i,k) = A(i,j) * B(j,k)", a, b, out_format=out_format, dtype=dtype)
