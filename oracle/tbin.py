"""Named-array container shared with the test oracle tooling (C twin: oracle/tbin.h).

Only used to move identical inputs between Python, the C oracle and the reference harness; it is not
part of the compute path.
"""
import struct

import numpy as np

_MAGIC = b"TB2BIN\0\0"
_DTYPES = {0: np.int32, 1: np.float32, 2: np.float64, 3: np.int64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def write(path, arrays):
    """arrays: dict name -> 1-D (or any-D, flattened C-order) numpy array of int32/float32/float64/int64."""
    with open(path, "wb") as f:
        f.write(_MAGIC)
        f.write(struct.pack("<II", len(arrays), 0))
        for name, arr in arrays.items():
            a = np.ascontiguousarray(arr).reshape(-1)
            code = _CODES[a.dtype]
            nm = name.encode()
            assert len(nm) < 24, name
            f.write(nm + b"\0" * (24 - len(nm)))
            f.write(struct.pack("<IIQ", code, 0, a.size))
            raw = a.tobytes()
            f.write(raw)
            f.write(b"\0" * ((-len(raw)) % 8))


def read(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == _MAGIC, "not a tbin file"
        n, _ = struct.unpack("<II", f.read(8))
        for _ in range(n):
            name = f.read(24).split(b"\0", 1)[0].decode()
            code, _, count = struct.unpack("<IIQ", f.read(16))
            dt = np.dtype(_DTYPES[code])
            nbytes = count * dt.itemsize
            out[name] = np.frombuffer(f.read(nbytes), dtype=dt).copy()
            f.read((-nbytes) % 8)
    return out
