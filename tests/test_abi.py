"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol include/taco_b200.h declares,
the statement classifier (module object) accepts exactly the hot-path patterns, and errors are reported the way the
header promises (non-zero return + message, never exit(), never a CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import taco_b200 as tb
from taco_b200 import _lib, formats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "taco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:_shim_)?taco_b200_\w+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_whole_surface():
    names = declared_symbols()
    for fam in _lib.FAMILIES:
        for ph in _lib.PHASES:
            assert f"taco_b200_{fam}_{ph}" in names and f"_shim_taco_b200_{fam}_{ph}" in names
    assert len(names) >= 2 * 3 * len(_lib.FAMILIES) + 20


@pytest.mark.parametrize("sym", declared_symbols())
def test_library_exports(sym):
    assert hasattr(_lib.lib, sym), f"{sym} declared in include/taco_b200.h but not exported by libtaco_b200.so"


def test_struct_layout_matches_reference():
    # /root/reference/include/taco/taco_tensor_t.h:13-23 on LP64: 4+pad, 8, 4+pad, 8, 8, 8, 8, 8, 4+pad = 72 bytes
    assert ctypes.sizeof(_lib.taco_tensor_t) == 72
    offs = {f[0]: getattr(_lib.taco_tensor_t, f[0]).offset for f in _lib.taco_tensor_t._fields_}
    assert offs == dict(order=0, dimensions=8, csize=16, mode_ordering=24, mode_types=32, indices=40, vals=48,
                        fill_value=56, vals_size=64)


CASES = [
    ("y(i) = A(i,j) * x(j)", "A:ds,x:d,y:d", "spmv"),
    ("a(r) = M(r,c) * v(c)", "M:ds", "spmv"),                       # renaming, unlisted tensors are dense
    ("C(i,k) = A(i,j) * B(j,k)", "A:ds,B:dd,C:dd", "spmm"),
    ("C(i,k) = A(i,j) * B(j,k)", "A:ds,B:dd,C:dd:1,0", "spmm"),      # reference GPU test's column-major result
    ("C(i,k) = A(i,j) * B(j,k)", "A:ss,B:dd,C:dd", "spmm_dcsr"),     # the reference's spmmDCSRGPU statement
    ("C(i,k) = A(i,j) * B(j,k)", "A:ds,B:ds,C:ds", "spgemm"),
    ("C(i,j) = A(i,j) + B(i,j)", "A:ds,B:ds,C:ds", "spadd"),
    ("A(i,j) = B(i,j) * C(i,k) * D(j,k)", "A:ds,B:ds,C:dd,D:dd", "sddmm"),
    ("A(i,k) = B(i,k) * C(i,j) * D(j,k)", "A:dd,B:ds,C:dd,D:dd", "sddmm_dense"),   # the reference's sddmmGPU statement
    ("A(i,j) = B(i,k,l) * C(k,j) * D(l,j)", "B:sss", "mttkrp"),
    ("A(i,j) = B(k,i,l) * C(k,j) * D(l,j)", "B:sss:1,0,2", "mttkrp"),       # parafac mode-J MTTKRP over the permuted storage
    ("A(i,j) = B(k,l,i) * C(k,j) * D(l,j)", "B:sss:2,0,1", "mttkrp"),       # mode-K
    ("A(i,j) = B(i,j,k) * c(k)", "B:sss", "ttv"),
    ("A(i,j,l) = B(i,j,k) * C(k,l)", "B:sss", "ttm"),
    ("a(i,j) = B(i,k,j,l) * c(k,l)", "B:dsdd", "bspmv"),                 # the reference's blocked SpMV statement
    ("C(i,j,m) = A(i,k,j,l) * B(k,l,m)", "A:dsdd,B:ddd,C:ddd", "bspmm"),
]


@pytest.mark.parametrize("expr,fmts,family", CASES)
def test_module_classifier(expr, fmts, family):
    for dt in (b"f32", b"f64"):
        m = _lib.lib.taco_b200_module_open(expr.encode(), fmts.encode(), dt)
        assert m, _lib.last_error()
        assert _lib.lib.taco_b200_module_family(m).decode() == family
        assert _lib.lib.taco_b200_module_num_args(m) == _lib.NARGS[family]
        for ph in _lib.PHASES:
            assert _lib.lib.taco_b200_module_get_func_ptr(m, ph.encode())
        assert _lib.lib.taco_b200_module_open(expr.encode(), fmts.encode(), dt) == m    # module cache hit


@pytest.mark.parametrize("expr,fmts", [
    ("y(i) = A(i,j) * x(j)", "A:dd"),                  # dense matvec: not a sparse hot-path kernel
    ("y(i) = A(j,i) * x(j)", "A:ds"),                  # transposed access
    ("C(i,j) = A(i,j) - B(i,j)", "A:ds,B:ds,C:ds"),
    ("C(i,k) = A(i,j) * B(j,k)", "A:ds,B:dd:1,0,C:dd"),
    ("a = B(i,j)", "B:ds"),
    ("A(i,j) = B(k,i,l) * C(k,j) * D(l,j)", "B:sss"),                       # mode-J MTTKRP needs B stored with mode 1 first
    ("A(i,j) = B(i,j,k) * c(k)", "B:sss:1,0,2"),
])
def test_module_refuses_everything_else(expr, fmts):
    assert not _lib.lib.taco_b200_module_open(expr.encode(), fmts.encode(), b"f64")
    assert "no CPU fallback" in _lib.last_error() or "cannot parse" in _lib.last_error()
    with pytest.raises(tb.TacoError):
        tb.Kernel(expr, fmts, "f64")


def test_unsupported_component_type():
    assert not _lib.lib.taco_b200_module_open(b"y(i) = A(i,j) * x(j)", b"A:ds", b"i32")


def test_partition_pos_host():
    pos = np.array([0, 2, 2, 5, 9, 9, 12], dtype=np.int32)
    assert tb.partition_pos(pos, 6, 3).tolist() == [0, 3, 4, 6]
    assert tb.partition_pos(pos, 6, 1).tolist() == [0, 6]
    b = tb.partition_pos(np.zeros(5, dtype=np.int32), 4, 4)          # empty tensor: all shards empty but covering
    assert b[0] == 0 and b[-1] == 4 and (np.diff(b) >= 0).all()


def test_no_cpu_fallback_without_device():
    """Without a GPU every compute entry point must fail loudly (the product path never runs on the CPU)."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    p, c, v = formats.csr_from_dense(np.eye(3))
    A = tb.makeCSR("A", [3, 3], p, c, v)
    x = tb.makeDense("x", [3], np.ones(3))
    y = tb.Tensor("y", [3], tb.Format([tb.dense]))
    k = tb.compile("y(i) = A(i,j) * x(j)", y, A, x)
    with pytest.raises(tb.TacoError) as ei:
        k(y, A, x)
    assert ei.value.code == 1 and "no CPU fallback" in str(ei.value)


def test_stub_source_for_compileSource_is_valid_c(tmp_path):
    """taco_b200_module_stub_source: the text handed to the reference's TensorBase::compileSource must define
    assemble/compute/evaluate with one taco_tensor_t* per tensor and compile as plain C (the reference JITs it with cc)."""
    import ctypes
    import subprocess
    from taco_b200 import _lib
    _lib.lib.taco_b200_module_stub_source.restype = ctypes.c_char_p
    _lib.lib.taco_b200_module_stub_source.argtypes = [ctypes.c_void_p]
    for expr, fm, fam, n in [("y(i) = A(i,j) * x(j)", "y:d,A:ds,x:d", "spmv", 3),
                             ("A(i,j) = B(i,k,l) * C(k,j) * D(l,j)", "A:dd,B:sss,C:dd,D:dd", "mttkrp", 4)]:
        m = _lib.lib.taco_b200_module_open(expr.encode(), fm.encode(), b"f64")
        assert m
        src = _lib.lib.taco_b200_module_stub_source(m).decode()
        for ph in ("assemble", "compute", "evaluate"):
            assert f"int {ph}(" + ", ".join(f"taco_tensor_t* t{a}" for a in range(n)) + ")" in src
            assert f"taco_b200_{fam}_{ph}" in src
        f = tmp_path / f"{fam}.c"
        f.write_text(src)
        subprocess.check_call(["/usr/bin/gcc", "-std=gnu99", "-Wall", "-Werror", "-shared", "-fPIC", str(f), "-o",
                               str(tmp_path / f"{fam}.so")])
