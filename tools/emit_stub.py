#!/usr/bin/env python
"""Print the C source that makes an UNMODIFIED taco run a statement on libtaco_b200 -- the text for
`TensorBase::compileSource()` / the CLI's `-read-source=`.

    python tools/emit_stub.py "y(i) = A(i,j) * x(j)" "y:d,A:ds,x:d" f64 > stub.c
    TACO_B200_LIB=taco_b200/lib/libtaco_b200.so taco "y(i) = A(i,j) * x(j)" -f=A:ds -i=A:a.mtx -i=x:x.tns -read-source=stub.c -verify
Needs no GPU: the classification happens in the library's module cache, kernels are only resolved when taco calls them."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taco_b200 import _lib  # noqa: E402


def stub_source(expr, formats, dtype="f64"):
    _lib.lib.taco_b200_module_stub_source.restype = ctypes.c_char_p
    _lib.lib.taco_b200_module_stub_source.argtypes = [ctypes.c_void_p]
    m = _lib.lib.taco_b200_module_open(expr.encode(), formats.encode(), dtype.encode())
    if not m:
        raise SystemExit("taco_b200: " + _lib.last_error())
    return _lib.lib.taco_b200_module_stub_source(m).decode()


if __name__ == "__main__":
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    sys.stdout.write(stub_source(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "f64"))
