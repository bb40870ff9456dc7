"""Shared test helpers: golden fixtures and the reference's own known-answer fixtures."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix):
    paths = sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "_*.npz")))
    assert paths, f"no golden fixtures for {prefix}"
    return [os.path.basename(p)[:-4] for p in paths]


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def tol(dtype):
    """north-star tolerances: 1e-12 relative for fp64, 1e-5 for fp32 (reduction reordering)"""
    return 1e-12 if np.dtype(dtype) == np.float64 else 1e-5


def assert_close(actual, expected, dtype, scale=None):
    actual, expected = np.asarray(actual), np.asarray(expected)
    assert actual.shape == expected.shape, (actual.shape, expected.shape)
    rt = tol(dtype)
    denom = np.maximum(np.abs(expected), 0 if scale is None else scale)
    err = np.abs(actual.astype(np.float64) - expected.astype(np.float64))
    bad = err > rt * np.maximum(denom, np.finfo(np.float64).tiny) + (0 if scale is not None else 0)
    # entries whose expected value is exactly 0 must match to within rt * scale (or exactly when no scale is given)
    assert not bad.any(), f"max rel err {np.max(err / np.maximum(denom, 1e-300)):.3e} > {rt:g} at {np.argmax(bad)}"


# ---- fixtures of the reference's own tests, /root/reference/test/test_tensors.cpp --------------------------
def d33a():   # :205-211
    d = np.zeros((3, 3)); d[0, 1] = 2; d[2, 0] = 3; d[2, 2] = 4
    return d


def d33b():   # :221-227
    d = np.zeros((3, 3)); d[0, 0] = 10; d[0, 1] = 20; d[2, 1] = 30
    return d


def d3b():    # :89-94
    return np.array([2.0, 0.0, 3.0])


def d34a():   # :237-244
    d = np.zeros((3, 4)); d[0, 0] = 2; d[0, 2] = 3; d[2, 0] = 4; d[2, 3] = 5
    return d


def d34b():   # :246-253
    d = np.zeros((3, 4)); d[0, 0] = 2; d[0, 3] = 3; d[2, 0] = 4; d[2, 2] = 5
    return d


def d233a():  # :296-305  -> (i, k, l, vals)
    c = [(0, 0, 0, 2), (0, 0, 1, 3), (0, 2, 2, 4), (1, 0, 1, 5), (1, 2, 0, 6), (1, 2, 2, 7)]
    i, k, l, v = map(np.array, zip(*c))
    return i, k, l, v.astype(np.float64)


def d333a():  # :328-339  -> (i, j, k, vals)
    c = [(0, 0, 0, 2), (0, 0, 1, 3), (0, 2, 2, 4), (1, 0, 1, 5), (1, 2, 0, 6), (1, 2, 2, 7), (2, 1, 2, 8), (2, 2, 1, 9)]
    i, j, k, v = map(np.array, zip(*c))
    return i, j, k, v.astype(np.float64)


def d3322a():  # :369-386  -> BCSR (pos, crd, blocks[nnzb,2,2]) of the order-4 fixture, dims (3,3,2,2)
    pos = np.array([0, 1, 1, 3], np.int32)
    crd = np.array([1, 0, 2], np.int32)
    blocks = np.array([[[2.1, 2.2], [2.3, 2.4]], [[3.1, 3.2], [3.3, 3.4]], [[4.1, 4.2], [4.3, 4.4]]])
    return pos, crd, blocks


def d32b():   # :358-367
    return np.array([[10.0, 11.0], [20.0, 21.0], [30.0, 31.0]])


def rua32_csr():
    """The CSR storage the reference expects after reading + packing test/data/rua_32.mtx (32 x 32, 126 entries):
    /root/reference/test/tests-api.cpp:261-300 (APIMatrixStorageTestData "rua_32.mtx", format CSR).  The file's values encode
    their coordinates, vals = 100 * (row + 1) + (col + 1), so the known answer is reproduced here from pos / crd alone."""
    pos = np.array([0, 6, 11, 17, 21, 25, 28, 33, 38, 45, 52, 57, 60, 62, 64, 67, 70, 73, 78, 81, 84, 87, 89, 93, 96, 101, 105,
                    109, 111, 116, 120, 123, 126], np.int32)
    crd = np.array([0, 1, 2, 3, 6, 25, 0, 1, 8, 20, 27, 1, 2, 5, 7, 8, 28, 2, 3, 4, 11, 2, 4, 22, 26, 0, 5, 15, 2, 6, 13, 20, 30,
                    0, 7, 11, 16, 26, 6, 8, 9, 12, 18, 22, 26, 0, 9, 10, 20, 22, 24, 26, 1, 10, 14, 17, 28, 5, 11, 23, 10, 12, 2,
                    13, 1, 14, 19, 3, 15, 21, 3, 15, 16, 5, 9, 17, 19, 29, 0, 18, 25, 7, 15, 19, 2, 20, 31, 10, 21, 1, 16, 20, 22,
                    11, 23, 25, 5, 14, 17, 23, 24, 12, 17, 21, 25, 4, 23, 25, 26, 8, 27, 2, 4, 26, 28, 31, 11, 16, 22, 29, 12, 13,
                    30, 23, 27, 31], np.int32)
    rows = np.repeat(np.arange(32), np.diff(pos))
    vals = (100.0 * (rows + 1) + (crd + 1)).astype(np.float64)
    return pos, crd, vals
