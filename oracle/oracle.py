"""ctypes front-end of the CPU oracle (oracle/taco_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; the
product package (taco_b200/) never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_I32P = ctypes.POINTER(ctypes.c_int32)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libtaco_oracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "taco_oracle.c")):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_get_max_threads.restype = ctypes.c_int
    return _LIB


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


def max_threads():
    return lib().oracle_get_max_threads()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i32(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a


def _sfx(dtype):
    dtype = np.dtype(dtype)
    assert dtype in (np.dtype(np.float32), np.dtype(np.float64)), dtype
    return "f32" if dtype == np.float32 else "f64"


def spmv(pos, crd, vals, x):
    pos, crd = _i32(pos), _i32(crd)
    n = pos.size - 1
    vals = np.ascontiguousarray(vals)
    x = np.ascontiguousarray(x, dtype=vals.dtype)
    y = np.empty(n, dtype=vals.dtype)
    getattr(lib(), "oracle_spmv_" + _sfx(vals.dtype))(ctypes.c_int32(n), _p(pos), _p(crd), _p(vals), _p(x), _p(y))
    return y


def spmm(pos, crd, vals, B):
    pos, crd = _i32(pos), _i32(crd)
    n = pos.size - 1
    vals = np.ascontiguousarray(vals)
    B = np.ascontiguousarray(B, dtype=vals.dtype)
    K = B.shape[1]
    C = np.empty((n, K), dtype=vals.dtype)
    getattr(lib(), "oracle_spmm_" + _sfx(vals.dtype))(ctypes.c_int32(n), ctypes.c_int32(K), _p(pos), _p(crd), _p(vals),
                                                     _p(B), _p(C))
    return C


def spmm_dcsr(n, pos1, crd1, pos2, crd2, vals, B):
    """C(i,k) = A(i,j) * B(j,k) with A = {Sparse,Sparse}; n = number of rows of A / C"""
    pos1, crd1, pos2, crd2 = _i32(pos1), _i32(crd1), _i32(pos2), _i32(crd2)
    vals = np.ascontiguousarray(vals)
    B = np.ascontiguousarray(B, dtype=vals.dtype)
    K = B.shape[1]
    C = np.empty((n, K), dtype=vals.dtype)
    getattr(lib(), "oracle_spmm_dcsr_" + _sfx(vals.dtype))(ctypes.c_int32(n), ctypes.c_int32(K), _p(pos1), _p(crd1), _p(pos2),
                                                          _p(crd2), _p(vals), _p(B), _p(C))
    return C


def sddmm_dense(pos, crd, bvals, C, D):
    """A(i,k) = B(i,k) * C(i,j) * D(j,k) with a dense result; C is (n, J), D is (J, m).  Returns A (n, m)."""
    pos, crd = _i32(pos), _i32(crd)
    bvals = np.ascontiguousarray(bvals)
    C = np.ascontiguousarray(C, dtype=bvals.dtype)
    D = np.ascontiguousarray(D, dtype=bvals.dtype)
    n, J = C.shape
    m = D.shape[1]
    A = np.empty((n, m), dtype=bvals.dtype)
    getattr(lib(), "oracle_sddmm_dense_" + _sfx(bvals.dtype))(ctypes.c_int32(n), ctypes.c_int32(m), ctypes.c_int32(J), _p(pos),
                                                              _p(crd), _p(bvals), _p(C), _p(D), _p(A))
    return A


def sddmm(pos, crd, bvals, C, D):
    """returns (A_pos, A_crd, A_vals)"""
    pos, crd = _i32(pos), _i32(crd)
    n = pos.size - 1
    bvals = np.ascontiguousarray(bvals)
    C = np.ascontiguousarray(C, dtype=bvals.dtype)
    D = np.ascontiguousarray(D, dtype=bvals.dtype)
    K = C.shape[1]
    apos = np.empty(n + 1, dtype=np.int32)
    crd_ptr = _I32P()
    lib().oracle_sddmm_assemble(ctypes.c_int32(n), _p(pos), _p(crd), _p(apos), ctypes.byref(crd_ptr))
    nnz = int(apos[n])
    acrd = np.ctypeslib.as_array(crd_ptr, shape=(max(nnz, 1),))[:nnz].copy()
    lib().oracle_free(crd_ptr)
    avals = np.empty(nnz, dtype=bvals.dtype)
    getattr(lib(), "oracle_sddmm_" + _sfx(bvals.dtype))(ctypes.c_int32(n), ctypes.c_int32(K), _p(pos), _p(crd),
                                                      _p(bvals), _p(C), _p(D), _p(avals))
    return apos, acrd, avals


def _csf_args(t):
    keys = ["B1_pos", "B1_crd", "B2_pos", "B2_crd", "B3_pos", "B3_crd"]
    arrs = [_i32(t[k]) for k in keys]
    return arrs


def mttkrp(t, C, D, I):
    vals = np.ascontiguousarray(t["B_vals"])
    C = np.ascontiguousarray(C, dtype=vals.dtype)
    D = np.ascontiguousarray(D, dtype=vals.dtype)
    R = C.shape[1]
    A = np.empty((I, R), dtype=vals.dtype)
    a = _csf_args(t)
    getattr(lib(), "oracle_mttkrp_" + _sfx(vals.dtype))(ctypes.c_int32(R), *[_p(x) for x in a], _p(vals), _p(C), _p(D),
                                                       ctypes.c_int32(I), _p(A))
    return A


def ttv(t, c, I, K):
    vals = np.ascontiguousarray(t["B_vals"])
    c = np.ascontiguousarray(c, dtype=vals.dtype)
    A = np.empty((I, K), dtype=vals.dtype)
    a = _csf_args(t)
    getattr(lib(), "oracle_ttv_" + _sfx(vals.dtype))(*[_p(x) for x in a], _p(vals), _p(c), ctypes.c_int32(I),
                                                    ctypes.c_int32(K), _p(A))
    return A


def ttm(t, C, I, K):
    vals = np.ascontiguousarray(t["B_vals"])
    C = np.ascontiguousarray(C, dtype=vals.dtype)
    R = C.shape[1]
    A = np.empty((I, K, R), dtype=vals.dtype)
    a = _csf_args(t)
    getattr(lib(), "oracle_ttm_" + _sfx(vals.dtype))(ctypes.c_int32(R), *[_p(x) for x in a], _p(vals), _p(C),
                                                    ctypes.c_int32(I), ctypes.c_int32(K), _p(A))
    return A


def bspmv(pos, crd, vals, c, br, bc):
    """a(i,j) = A(i,k,j,l) * c(k,l); vals = blocks [nnzb, br, bc], c = (Nb, bc); returns a (Mb, br)"""
    pos, crd = _i32(pos), _i32(crd)
    Mb = pos.size - 1
    vals = np.ascontiguousarray(vals)
    c = np.ascontiguousarray(c, dtype=vals.dtype)
    a = np.empty((Mb, br), dtype=vals.dtype)
    getattr(lib(), "oracle_bspmv_" + _sfx(vals.dtype))(ctypes.c_int32(Mb), ctypes.c_int32(br), ctypes.c_int32(bc),
                                                      _p(pos), _p(crd), _p(vals), _p(c), _p(a))
    return a


def bspmm(pos, crd, vals, B, br, bc):
    """C(i,j,m) = A(i,k,j,l) * B(k,l,m); vals = blocks [nnzb, br, bc], B = (Nb*bc, K); returns C (Mb*br, K)"""
    pos, crd = _i32(pos), _i32(crd)
    Mb = pos.size - 1
    vals = np.ascontiguousarray(vals)
    B = np.ascontiguousarray(B, dtype=vals.dtype)
    K = B.shape[-1]
    C = np.empty((Mb * br, K), dtype=vals.dtype)
    getattr(lib(), "oracle_bspmm_" + _sfx(vals.dtype))(ctypes.c_int32(Mb), ctypes.c_int32(br), ctypes.c_int32(bc),
                                                      ctypes.c_int32(K), _p(pos), _p(crd), _p(vals), _p(B), _p(C))
    return C


def _sparse_out(assemble_fn, compute_name, n, head_args, struct_args, Aval, Bval):
    cpos = np.empty(n + 1, dtype=np.int32)
    crd_ptr = _I32P()
    assemble_fn(*head_args, *[_p(x) for x in struct_args], _p(cpos), ctypes.byref(crd_ptr))
    nnz = int(cpos[n])
    ccrd = np.ctypeslib.as_array(crd_ptr, shape=(max(nnz, 1),))[:nnz].copy()
    lib().oracle_free(crd_ptr)
    cvals = np.empty(nnz, dtype=Aval.dtype)
    Apos, Acrd, Bpos, Bcrd = struct_args
    getattr(lib(), compute_name + _sfx(Aval.dtype))(*head_args, _p(Apos), _p(Acrd), _p(Aval), _p(Bpos), _p(Bcrd),
                                                   _p(Bval), _p(cpos), _p(cvals))
    return cpos, ccrd, cvals


def spadd(Apos, Acrd, Aval, Bpos, Bcrd, Bval):
    """returns (C_pos, C_crd, C_vals)"""
    Apos, Acrd, Bpos, Bcrd = _i32(Apos), _i32(Acrd), _i32(Bpos), _i32(Bcrd)
    Aval = np.ascontiguousarray(Aval)
    Bval = np.ascontiguousarray(Bval, dtype=Aval.dtype)
    n = Apos.size - 1
    return _sparse_out(lib().oracle_spadd_assemble, "oracle_spadd_compute_", n, [ctypes.c_int32(n)],
                       [Apos, Acrd, Bpos, Bcrd], Aval, Bval)


def spgemm(Apos, Acrd, Aval, Bpos, Bcrd, Bval, ncols):
    """returns (C_pos, C_crd, C_vals)"""
    Apos, Acrd, Bpos, Bcrd = _i32(Apos), _i32(Acrd), _i32(Bpos), _i32(Bcrd)
    Aval = np.ascontiguousarray(Aval)
    Bval = np.ascontiguousarray(Bval, dtype=Aval.dtype)
    n = Apos.size - 1
    return _sparse_out(lib().oracle_spgemm_assemble, "oracle_spgemm_compute_", n,
                       [ctypes.c_int32(n), ctypes.c_int32(ncols)], [Apos, Acrd, Bpos, Bcrd], Aval, Bval)


def pack(kind, dims, coords, vals):
    """TensorBase::pack() restated (src/tensor.cpp:295-463): sort the coordinates lexicographically, add the values of
    equal coordinates, build the level arrays of {Dense,Compressed} ("csr"), {Compressed,Compressed} ("dcsr") or
    {Compressed x3} ("csf3") the way the generated `pack` helper appends them.  The sort here is stable, so duplicates are
    added in insertion order (the reference's qsort leaves that order unspecified).  Returns a dict of level arrays named
    like oracle/ref_harness.cpp's pack output: A<level>_pos / A<level>_crd (compressed levels only) and A_vals."""
    coords = [np.asarray(c, dtype=np.int64) for c in coords]
    if kind == "csc":          # {Dense,Compressed} with mode ordering {1,0}: CSR of the transposed coordinates
        kind, dims, coords = "csr", [dims[1], dims[0]], [coords[1], coords[0]]
    vals = np.ascontiguousarray(vals)
    n = vals.size
    order = len(coords)
    perm = np.lexsort(coords[::-1]) if n else np.zeros(0, np.int64)          # stable, first mode most significant
    sc = [c[perm] for c in coords]
    sv = vals[perm]
    new = np.ones(n, bool)
    if n:
        new[1:] = np.logical_or.reduce([c[1:] != c[:-1] for c in sc])
    start = np.flatnonzero(new)
    out_vals = np.zeros(start.size, dtype=vals.dtype)
    ends = np.append(start[1:], n)
    for u, (a, b) in enumerate(zip(start, ends)):                              # sequential adds, insertion order
        acc = sv[a]
        for e in range(a + 1, b):
            acc = acc + sv[e]
        out_vals[u] = acc
    uc = [c[start] for c in sc]                                                # distinct coordinates, sorted
    res = {"A_vals": out_vals}
    nu = start.size
    if kind == "csr":
        pos = np.zeros(dims[0] + 1, np.int64)
        np.add.at(pos, uc[0] + 1, 1)
        res.update(A2_pos=np.cumsum(pos).astype(np.int32), A2_crd=uc[1].astype(np.int32))
        return res
    # compressed levels from the top: nodes of level l = distinct prefixes (c0..cl)
    parent_ids = np.zeros(nu, np.int64)
    nparents = 1
    for l in range(order):
        head = np.ones(nu, bool)
        if nu:
            head[1:] = np.logical_or.reduce([c[1:] != c[:-1] for c in uc[: l + 1]])
        node = np.cumsum(head) - 1                                             # node index of every distinct entry
        nnodes = int(head.sum())
        pos = np.zeros(nparents + 1, np.int64)
        np.add.at(pos, parent_ids[head] + 1, 1)
        res[f"A{l + 1}_pos"] = np.cumsum(pos).astype(np.int32)
        res[f"A{l + 1}_crd"] = uc[l][head].astype(np.int32)
        parent_ids, nparents = node, nnodes
    return res


# ---- file readers restated (test infrastructure): /root/reference/src/storage/file_io_mtx.cpp:39-150, file_io_tns.cpp:39-96 ----
def read_mtx(path):
    """Matrix Market coordinate file -> (dims, [coords per mode], vals) in insertion order (0-based), as readMTX + readSparse
    insert them: header, '%' comments, size line `d1 d2 [..] nnz`, then nnz entries with 1-based indices and a strtod value;
    symmetric files insert the transposed entry right after every off-diagonal one."""
    with open(path, "rb") as fh:
        lines = fh.read().decode("ascii", "replace").split("\n")
    head = lines[0].split()
    assert head[0] == "%%MatrixMarket" and head[1] in ("matrix", "tensor") and head[2] == "coordinate" and head[3] == "real"
    symm = head[4] == "symmetric"
    q = 1
    while lines[q].split() and lines[q].split()[0].startswith("%"):
        q += 1
    size = [int(t) for t in lines[q].split()]
    dims, nnz = size[:-1], size[-1]
    coords = [[] for _ in dims]
    vals = []
    taken = 0
    for line in lines[q + 1:]:
        t = line.split()
        if not t or taken >= nnz:
            continue
        c = [int(x) - 1 for x in t[:len(dims)]]
        v = float(t[len(dims)])               # Python's float() is correctly rounded, like strtod
        for m, x in enumerate(c):
            coords[m].append(x)
        vals.append(v)
        taken += 1
        if symm and c[0] != c[-1]:
            for m, x in enumerate(reversed(c)):
                coords[m].append(x)
            vals.append(v)
    return dims, [np.array(c, np.int64) for c in coords], np.array(vals, np.float64)


def read_tns(path):
    """FROSTT file -> (dims, coords, vals): order from the first line, dimensions = largest coordinate per mode (readTNS)"""
    coords, vals = None, []
    with open(path, "rb") as fh:
        for line in fh.read().decode("ascii", "replace").split("\n"):
            t = line.split()
            if not t:
                continue
            if coords is None:
                coords = [[] for _ in t[:-1]]
            for m in range(len(coords)):
                coords[m].append(int(t[m]) - 1)
            vals.append(float(t[len(coords)]))
    dims = [int(max(c)) + 1 for c in coords]
    return dims, [np.array(c, np.int64) for c in coords], np.array(vals, np.float64)
