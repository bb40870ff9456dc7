"""ctypes binding of libtaco_b200.so (the C-ABI product library, include/taco_b200.h).

There is no fallback of any kind: if the shared library is missing this module raises at import, and every
entry point fails (TacoError) when no sm_100 device is usable.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TACO_B200_LIB (the variable the compileSource stub and the patched taco use) may point at another build of the library
LIB_PATH = os.environ.get("TACO_B200_LIB") or os.path.join(_HERE, "lib", "libtaco_b200.so")


class TacoError(RuntimeError):
    """Mirror of taco::TacoException for failures reported by the library (non-zero return code)."""

    def __init__(self, code, message):
        super().__init__(f"taco_b200 error {code}: {message}")
        self.code = code


class taco_tensor_t(ctypes.Structure):
    # layout of /root/reference/include/taco/taco_tensor_t.h:13-23 (restated in include/taco_b200.h)
    _fields_ = [
        ("order", ctypes.c_int32),
        ("dimensions", ctypes.POINTER(ctypes.c_int32)),
        ("csize", ctypes.c_int32),
        ("mode_ordering", ctypes.POINTER(ctypes.c_int32)),
        ("mode_types", ctypes.POINTER(ctypes.c_int32)),
        ("indices", ctypes.POINTER(ctypes.POINTER(ctypes.c_void_p))),
        ("vals", ctypes.c_void_p),
        ("fill_value", ctypes.c_void_p),
        ("vals_size", ctypes.c_int32),
    ]


def build(verbose=False):
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    import subprocess
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    subprocess.check_call(cmd if verbose else cmd + ["-s"])


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(make -C taco_b200/csrc). taco_b200 has no Python/CPU fallback path.")

lib = ctypes.CDLL(LIB_PATH)


class _Optional:
    """argtypes of entry points an older build of the library (A/B runs through TACO_B200_LIB) may not export yet"""
    argtypes = restype = None


def _sym(name):
    return getattr(lib, name) if hasattr(lib, name) else _Optional()


FAMILIES = ("spmv", "spmm", "spmm_dcsr", "sddmm", "sddmm_dense", "mttkrp", "ttv", "ttm", "spadd", "spgemm", "bspmv", "bspmm")
NARGS = {"spmv": 3, "spmm": 3, "spmm_dcsr": 3, "sddmm": 4, "sddmm_dense": 4, "mttkrp": 4, "ttv": 3, "ttm": 3, "spadd": 3, "spgemm": 3, "bspmv": 3,
         "bspmm": 3}
PHASES = ("assemble", "compute", "evaluate")

_TP = ctypes.POINTER(taco_tensor_t)
lib.taco_b200_last_error.restype = ctypes.c_char_p
lib.taco_b200_version.restype = ctypes.c_char_p
lib.taco_b200_get_stream.restype = ctypes.c_void_p
lib.taco_b200_set_stream.argtypes = [ctypes.c_void_p]
lib.taco_b200_host_alloc.restype = ctypes.c_void_p
lib.taco_b200_host_alloc.argtypes = [ctypes.c_size_t]
lib.taco_b200_host_free.argtypes = [ctypes.c_void_p]
lib.taco_b200_device_alloc.restype = ctypes.c_void_p
lib.taco_b200_device_alloc.argtypes = [ctypes.c_size_t]
lib.taco_b200_free.argtypes = [ctypes.c_void_p]
_sym("taco_b200_set_result_multicast").argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
_sym("taco_b200_set_result_peers").argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
lib.taco_b200_make_resident.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
lib.taco_b200_invalidate.argtypes = [ctypes.c_void_p]
lib.taco_b200_module_open.restype = ctypes.c_void_p
lib.taco_b200_module_open.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
_sym("taco_b200_module_open_args").restype = ctypes.c_void_p
_sym("taco_b200_module_open_args").argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
lib.taco_b200_module_family.restype = ctypes.c_char_p
lib.taco_b200_module_family.argtypes = [ctypes.c_void_p]
lib.taco_b200_module_num_args.argtypes = [ctypes.c_void_p]
lib.taco_b200_module_call_packed.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
lib.taco_b200_module_get_func_ptr.restype = ctypes.c_void_p
lib.taco_b200_module_get_func_ptr.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
lib.taco_b200_pack.argtypes = [_TP, _TP]
lib._shim_taco_b200_pack.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
_sym("taco_b200_read").argtypes = [ctypes.c_char_p, _TP]
_sym("_shim_taco_b200_read").argtypes = [ctypes.POINTER(ctypes.c_void_p)]
lib.taco_b200_partition_pos.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
for _f in FAMILIES:
    for _p in PHASES:
        getattr(lib, f"taco_b200_{_f}_{_p}").argtypes = [_TP] * NARGS[_f]
        getattr(lib, f"_shim_taco_b200_{_f}_{_p}").argtypes = [ctypes.POINTER(ctypes.c_void_p)]


def check(rc):
    if rc != 0:
        raise TacoError(rc, (lib.taco_b200_last_error() or b"").decode())
    return rc


def last_error():
    return (lib.taco_b200_last_error() or b"").decode()
