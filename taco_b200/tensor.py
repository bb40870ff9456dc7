"""Host-side mirror of the reference's tensor / kernel interface for the GPU hot path.

Names and meaning follow the reference's C++ API so that tests read like the reference's own:
  Format / ModeFormat (dense, compressed)       /root/reference/include/taco/format.h:21-214
  makeCSR (zero-copy attach of pos/crd/vals)     /root/reference/include/taco/tensor.h:774-797
  Index/ModeIndex attach for CSF                 /root/reference/include/taco/storage/index.h:20-75
  compile(stmt) -> Kernel{assemble,compute,()}   /root/reference/src/index_notation/kernel.cpp:83-128
  TensorStorage -> taco_tensor_t*                /root/reference/src/storage/storage.cpp:127-177
Arrays may be numpy arrays (host; pageable, or pinned via `pinned_empty`) or torch CUDA tensors (device
resident, used in place).  Everything is executed by libtaco_b200.so through the taco_tensor_t C ABI.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import TacoError, lib, check, taco_tensor_t

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

dense = "d"
compressed = "s"
Dense = dense
Sparse = compressed

_libc = ctypes.CDLL(None)
_libc.free.argtypes = [ctypes.c_void_p]


class Format:
    """Format({Dense, Sparse}, ordering) -- level formats in storage order plus the mode ordering."""

    def __init__(self, levels, ordering=None):
        self.levels = "".join(levels)
        self.ordering = list(ordering) if ordering is not None else list(range(len(self.levels)))
        assert sorted(self.ordering) == list(range(len(self.levels)))

    def spec(self):
        s = self.levels
        if self.ordering != list(range(len(self.levels))):
            s += ":" + ",".join(map(str, self.ordering))
        return s

    def __eq__(self, o):
        return isinstance(o, Format) and (self.levels, self.ordering) == (o.levels, o.ordering)


CSR = Format([dense, compressed])
CSF3 = Format([compressed, compressed, compressed])
DCSR = Format([compressed, compressed])               # doubly compressed rows: only rows with nonzeros are stored
BCSR = Format([dense, compressed, dense, dense])      # (block row, block column, row in block, column in block)


def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return a.data_ptr()
    return a.ctypes.data


def _np_dtype(dt):
    dt = np.dtype(dt)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TacoError(2, f"component type {dt} is not supported on the GPU path (float32/float64)")
    return dt


class _DeviceArray:
    """Wraps a raw device pointer returned by the library so torch can adopt it without a copy."""

    def __init__(self, ptr, count, dtype):
        self.ptr, self.count, self.dtype = ptr, count, np.dtype(dtype)
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": self.dtype.str, "data": (ptr or 0, False),
                                         "version": 2}

    def __del__(self):
        if self.ptr:
            lib.taco_b200_free(ctypes.c_void_p(self.ptr))
            self.ptr = None


class _HostArray:
    """malloc()ed result array handed back by assemble (the reference frees these with free(), Array::Free)."""

    def __init__(self, ptr, count, dtype):
        self.ptr = ptr
        self.array = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(np.ctypeslib.as_ctypes_type(dtype))),
                                           shape=(max(count, 1),))[:count]

    def __del__(self):
        if self.ptr:
            _libc.free(ctypes.c_void_p(self.ptr))
            self.ptr = None


class Tensor:
    """A named tensor bound to a taco_tensor_t.  Operands are attached zero-copy; results get their arrays from
    assemble()."""

    def __init__(self, name, dims, fmt, dtype=np.float64):
        self.name = name
        self.dims = [int(d) for d in dims]
        self.format = fmt if isinstance(fmt, Format) else Format(fmt)
        self.dtype = _np_dtype(dtype)
        order = len(self.dims)
        assert order == len(self.format.levels)
        self._keep = []          # python objects that own memory referenced by the struct
        self._owned = {}         # result arrays allocated by the library
        self.arrays = {}         # (level, k) -> array ; "vals" -> array
        t = taco_tensor_t()
        t.order = order
        self._dims = (ctypes.c_int32 * max(order, 1))(*self.dims)
        self._ordering = (ctypes.c_int32 * max(order, 1))(*self.format.ordering)
        self._types = (ctypes.c_int32 * max(order, 1))(*[0 if c == dense else 1 for c in self.format.levels])
        t.dimensions = ctypes.cast(self._dims, ctypes.POINTER(ctypes.c_int32))
        t.mode_ordering = ctypes.cast(self._ordering, ctypes.POINTER(ctypes.c_int32))
        t.mode_types = ctypes.cast(self._types, ctypes.POINTER(ctypes.c_int32))
        t.csize = self.dtype.itemsize * 8
        self._fill = (ctypes.c_uint8 * 8)()
        t.fill_value = ctypes.cast(self._fill, ctypes.c_void_p)
        self._level_ptrs = []
        self._dimsz = []
        idx = (ctypes.POINTER(ctypes.c_void_p) * max(order, 1))()
        for l, c in enumerate(self.format.levels):
            arr = (ctypes.c_void_p * 2)()
            if c == dense:
                # dense level: indices[l][0] -> int32[1]{dimension of the mode stored at this level}
                dz = (ctypes.c_int32 * 1)(self.dims[self.format.ordering[l]])
                self._dimsz.append(dz)
                arr[0] = ctypes.cast(dz, ctypes.c_void_p)
            self._level_ptrs.append(arr)
            idx[l] = ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p))
        self._idx = idx
        t.indices = ctypes.cast(idx, ctypes.POINTER(ctypes.POINTER(ctypes.c_void_p)))
        t.vals = None
        t.vals_size = 0
        self.ct = t

    # ---- attaching operand arrays --------------------------------------------------------------------
    def set_level(self, level, pos, crd):
        for k, a in enumerate((pos, crd)):
            if a is not None:
                self._check_index(a)
                self._level_ptrs[level][k] = _ptr(a)
                self.arrays[(level, k)] = a
        return self

    def set_vals(self, vals, count=None):
        if _is_torch(vals):
            assert vals.is_contiguous()
            want = torch.float32 if self.dtype == np.float32 else torch.float64
            assert vals.dtype == want, (vals.dtype, want)
        else:
            assert vals.flags["C_CONTIGUOUS"] and vals.dtype == self.dtype, (vals.dtype, self.dtype)
        self.ct.vals = _ptr(vals)
        self.arrays["vals"] = vals
        n = int(vals.numel() if _is_torch(vals) else vals.size) if count is None else count
        # convention for device-resident tensors (include/taco_b200.h): vals_size = number of stored values
        self.ct.vals_size = n if n < 2**31 else 0
        return self

    @staticmethod
    def _check_index(a):
        if _is_torch(a):
            assert a.dtype == torch.int32 and a.is_contiguous()
        else:
            assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]

    @property
    def ptr(self):
        return ctypes.pointer(self.ct)

    # ---- reading results ------------------------------------------------------------------------------
    def _adopt(self, key, ptr, count, dtype):
        """Take ownership of an array the library allocated in assemble()."""
        if not ptr:
            return None
        prev = self._owned.get(key)
        if prev is not None and prev.ptr == ptr:
            return prev
        if _lib.lib.taco_b200_get_result_space() == 1:
            obj = _DeviceArray(ptr, count, dtype)
        else:
            obj = _HostArray(ptr, count, dtype)
        self._owned[key] = obj
        return obj

    def _as_user_array(self, obj):
        if obj is None:
            return None
        if isinstance(obj, _DeviceArray):
            return torch.as_tensor(obj, device="cuda") if obj.count else torch.empty(0, device="cuda")
        return obj.array

    def adopt_results(self, nnz=None):
        """After assemble(): wrap result arrays (pos/crd/vals) so they are freed with the tensor."""
        size = 1
        for l, c in enumerate(self.format.levels):
            if c == dense:
                size *= self.dims[self.format.ordering[l]]
            else:
                pos_ptr, crd_ptr = self._level_ptrs[l][0], self._level_ptrs[l][1]
                pos = self._adopt((l, 0), pos_ptr, size + 1, np.int32)
                if compressed not in self.format.levels[l + 1:]:          # last compressed level: one node per stored value
                    child = int(self.ct.vals_size) if nnz is None else nnz
                else:           # an inner compressed level (DCSR / CSF results of pack()): its node count is pos[parent size]
                    if isinstance(pos, _DeviceArray):
                        check(lib.taco_b200_synchronize())      # the library's stream wrote it; torch reads on its own stream
                    child = int(self.to_numpy(self._as_user_array(pos))[size])
                crd = self._adopt((l, 1), crd_ptr, child, np.int32)
                self.arrays[(l, 0)] = self._as_user_array(pos)
                self.arrays[(l, 1)] = self._as_user_array(crd)
                size = child
        vals = self._adopt("vals", self.ct.vals, size, self.dtype)
        self.arrays["vals"] = self._as_user_array(vals)
        return self

    def vals(self):
        return self.arrays.get("vals")

    def level(self, l):
        return self.arrays.get((l, 0)), self.arrays.get((l, 1))

    def to_numpy(self, a):
        if _is_torch(a):
            return a.detach().cpu().numpy()
        return np.asarray(a)


# ---- factory functions (names follow the reference) -------------------------------------------------------
def makeCSR(name, dims, pos, crd, vals):
    dt = np.float32 if (_is_torch(vals) and vals.dtype == torch.float32) or (not _is_torch(vals) and vals.dtype == np.float32) else np.float64
    t = Tensor(name, dims, CSR, dt)
    t.set_level(1, pos, crd)
    t.set_vals(vals)
    return t


def makeDense(name, dims, vals, ordering=None):
    dt = np.float32 if (_is_torch(vals) and vals.dtype == torch.float32) or (not _is_torch(vals) and vals.dtype == np.float32) else np.float64
    t = Tensor(name, dims, Format([dense] * len(dims), ordering), dt)
    t.set_vals(vals)
    return t


def makeCSF3(name, dims, arrays, ordering=None):
    """arrays: dict with B1_pos,B1_crd,B2_pos,B2_crd,B3_pos,B3_crd,B_vals (taco_b200.formats.coo_to_csf3 layout).
    `dims` are indexed by mode; with a mode `ordering` (e.g. [1, 0, 2]) level l of the arrays stores mode ordering[l]."""
    vals = arrays["B_vals"]
    dt = np.float32 if (_is_torch(vals) and vals.dtype == torch.float32) or (not _is_torch(vals) and vals.dtype == np.float32) else np.float64
    t = Tensor(name, dims, CSF3 if ordering is None else Format([compressed] * 3, ordering), dt)
    for l in range(3):
        t.set_level(l, arrays[f"B{l + 1}_pos"], arrays[f"B{l + 1}_crd"])
    t.set_vals(vals)
    return t


def makeDCSR(name, dims, arrays):
    """arrays: dict with A1_pos,A1_crd,A2_pos,A2_crd,A_vals (taco_b200.formats.dcsr_from_dense layout), zero-copy."""
    vals = arrays["A_vals"]
    dt = np.float32 if (_is_torch(vals) and vals.dtype == torch.float32) or (not _is_torch(vals) and vals.dtype == np.float32) else np.float64
    t = Tensor(name, dims, DCSR, dt)
    for l in range(2):
        t.set_level(l, arrays[f"A{l + 1}_pos"], arrays[f"A{l + 1}_crd"])
    t.set_vals(vals)
    return t


def pack(name, dims, fmt, coords, vals):
    """COO -> a packed tensor on the GPU: the mirror of `Tensor::insert()` x n + `pack()` (src/tensor.cpp:295-463).
    coords = one int32 array per mode (any order, duplicates are added), vals = the components; numpy or CUDA tensors.
    Returns a Tensor of format `fmt` (CSR, DCSR or CSF3) whose level arrays live in the current result space."""
    dt = np.float32 if (_is_torch(vals) and vals.dtype == torch.float32) or (not _is_torch(vals) and vals.dtype == np.float32) else np.float64
    order = len(dims)
    n = int(vals.numel() if _is_torch(vals) else vals.size)
    fmt = fmt if isinstance(fmt, Format) else Format(fmt)
    coo = Tensor(name + "_coo", dims, Format([compressed] * order, fmt.ordering), dt)
    coo.set_level(0, np.array([0, n], np.int32), coords[fmt.ordering[0]])       # level l holds mode ordering[l]
    for l in range(1, order):
        coo.set_level(l, None, coords[fmt.ordering[l]])
    coo.set_vals(vals)
    t = Tensor(name, dims, fmt, dt)
    check(lib.taco_b200_pack(t.ptr, coo.ptr))
    t.adopt_results()
    return t


def read(path, fmt, dtype=np.float64, dims=None, name="A"):
    """taco::read(filename, format) for .mtx / .ttx (Matrix Market coordinate) and .tns (FROSTT) files
    (/root/reference/src/tensor.cpp read(), src/storage/file_io_mtx.cpp, file_io_tns.cpp): the file is parsed and packed on
    the device (taco_b200_read).  `dims` = None takes the dimensions from the file, as the reference does."""
    fmt = fmt if isinstance(fmt, Format) else Format(fmt)
    order = len(fmt.levels)
    t = Tensor(name, list(dims) if dims is not None else [0] * order, fmt, dtype)
    check(lib.taco_b200_read(str(path).encode(), t.ptr))
    t.dims = [int(t._dims[m]) for m in range(order)]
    t.adopt_results()
    return t


def makeBCSR(name, dims, pos, crd, vals):
    """dims = [Mb, Nb, br, bc]; pos/crd over the block columns, vals = [stored blocks * br * bc] (zero-copy attach)."""
    dt = np.float32 if (_is_torch(vals) and vals.dtype == torch.float32) or (not _is_torch(vals) and vals.dtype == np.float32) else np.float64
    t = Tensor(name, dims, BCSR, dt)
    t.set_level(1, pos, crd)
    t.set_vals(vals)
    return t


def pinned_empty(shape, dtype):
    """numpy array backed by pinned host memory from the library (fast H2D/D2H staging)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    ptr = lib.taco_b200_host_alloc(max(count, 1) * dtype.itemsize)
    if not ptr:
        raise TacoError(5, _lib.last_error())
    buf = (ctypes.c_uint8 * (max(count, 1) * dtype.itemsize)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _PINNED[arr.ctypes.data] = ptr
    return arr


_PINNED = {}


def pinned_free(arr):
    ptr = _PINNED.pop(arr.ctypes.data, None)
    if ptr:
        lib.taco_b200_host_free(ctypes.c_void_p(ptr))


# ---- Kernel ------------------------------------------------------------------------------------------------
class Kernel:
    """compile("y(i) = A(i,j) * x(j)", y, A, x): the GPU-path equivalent of taco::compile(stmt) -> Kernel.

    Tensors are passed in the reference's argument order: results first, then operands in order of first
    appearance (src/tensor.cpp:778-806)."""

    def __init__(self, expr, formats, dtype):
        self.expr = expr
        self.mod = lib.taco_b200_module_open(expr.encode(), formats.encode(), dtype.encode())
        if not self.mod:
            raise TacoError(4, _lib.last_error())
        self.family = lib.taco_b200_module_family(self.mod).decode()
        self.nargs = lib.taco_b200_module_num_args(self.mod)

    def _call(self, phase, tensors):
        if len(tensors) != self.nargs:
            raise TacoError(3, f"{self.family} takes {self.nargs} tensors, got {len(tensors)}")
        pack = (ctypes.c_void_p * self.nargs)(*[ctypes.cast(t.ptr, ctypes.c_void_p) for t in tensors])
        check(lib.taco_b200_module_call_packed(self.mod, phase.encode(), pack))

    def assemble(self, *tensors):
        self._call("assemble", tensors)
        tensors[0].adopt_results()
        return True

    def compute(self, *tensors):
        self._call("compute", tensors)
        return True

    def __call__(self, *tensors):
        self._call("evaluate", tensors)
        tensors[0].adopt_results()
        return True

    evaluate = __call__


def compile(expr, *tensors):  # noqa: A001  (name follows taco::compile)
    fm = ",".join(f"{t.name}:{t.format.spec()}" for t in tensors)
    dts = {t.dtype for t in tensors}
    if len(dts) != 1:
        raise TacoError(2, "mixed component types")
    return Kernel(expr, fm, "f32" if dts.pop() == np.float32 else "f64")


# ---- runtime controls ---------------------------------------------------------------------------------------
def set_result_space(space):
    check(lib.taco_b200_set_result_space({"host": 0, "device": 1}[space]))


def use_torch_stream():
    """Enqueue all library work on torch's current CUDA stream (so torch.cuda.Event brackets it)."""
    h = torch.cuda.current_stream().cuda_stream
    # torch's default stream is the legacy NULL stream (handle 0); the C ABI keeps NULL for "private stream", so
    # pass the explicit cudaStreamLegacy handle (0x1) instead
    check(lib.taco_b200_set_stream(ctypes.c_void_p(h if h else 1)))


def set_result_multicast(local_ptr, multicast_ptr, nbytes):
    """dense results inside [local_ptr, local_ptr + nbytes) are stored through the NVLink multicast mapping (None clears)"""
    check(lib.taco_b200_set_result_multicast(ctypes.c_void_p(local_ptr or 0), ctypes.c_void_p(multicast_ptr or 0), int(nbytes or 0)))


def set_result_peers(local_ptr, peer_ptrs, nbytes):
    """dense results inside [local_ptr, local_ptr + nbytes) are also stored, from inside the kernels, into the same window of
    every peer GPU (`peer_ptrs`: the peers' windows as mapped into this process, at most 7; None / empty clears)"""
    peers = [int(p) for p in (peer_ptrs or [])]
    arr = (ctypes.c_void_p * max(len(peers), 1))(*peers)
    check(lib.taco_b200_set_result_peers(ctypes.c_void_p(local_ptr or 0), int(nbytes or 0), len(peers), arr))


def synchronize():
    check(lib.taco_b200_synchronize())


def launch_count():
    return lib.taco_b200_launch_count()


def partition_pos(pos, parent_size, parts):
    bounds = np.zeros(parts + 1, dtype=np.int32)
    check(lib.taco_b200_partition_pos(ctypes.c_void_p(_ptr(pos)), int(parent_size), int(parts),
                                      ctypes.c_void_p(bounds.ctypes.data)))
    return bounds
