"""Device-side file -> tensor (`-m gpu`): taco_b200_read parses .mtx / .tns bytes on the GPU and packs them; the result must
equal the reference's reader + pack() restated in oracle/oracle.py (read_mtx / read_tns + pack) -- structure bit-exact,
values bit-identical to strtod, whatever the numeral looks like."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "oracle"), HERE]
import oracle  # noqa: E402
import helpers as H  # noqa: E402
import gpu_util as G  # noqa: E402
import taco_b200 as tb  # noqa: E402

pytestmark = pytest.mark.gpu

NUMERALS = ["1", "-2.25", "3.", ".5", "1.5e-3", "6.02214076E+23", "1e-30", "-0.000123456789012345678", "12345678901234567890",
            "0.1", "1.7976931348623157e308", "4.9e-324", "9007199254740993", "0.30000000000000004", "123456.789e-7", "+7", "0", "-0.0",
            "2.2250738585072014e-308", "1234567890123456789e-19", "8.5", "1e22", "1e23", "3.141592653589793238462643383279"]


def _check(t, fmt_kind, dims, coords, vals, dtype):
    want = oracle.pack(fmt_kind, dims, coords, vals.astype(dtype))
    assert t.dims == list(dims)
    levels = {"csr": [1], "csc": [1], "dcsr": [0, 1], "csf3": [0, 1, 2]}[fmt_kind]
    for l in levels:
        pos, crd = t.level(l)
        assert np.array_equal(G.to_host(pos), want[f"A{l + 1}_pos"]), f"pos of level {l} must be bit-exact"
        assert np.array_equal(G.to_host(crd), want[f"A{l + 1}_crd"]), f"crd of level {l} must be bit-exact"
    got = G.to_host(t.vals())
    assert got.dtype == np.dtype(dtype) and np.array_equal(got.view(np.uint8), want["A_vals"].view(np.uint8)), "values must equal strtod's bits"


def _write_mtx(path, dims, entries, symmetric=False, crlf=False, kind="matrix"):
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as fh:
        fh.write(f"%%MatrixMarket {kind} coordinate real {'symmetric' if symmetric else 'general'}{nl}")
        fh.write(f"%-----{nl}% a comment line{nl}")
        fh.write(" ".join(str(d) for d in dims) + f" {len(entries)}{nl}")
        for e in entries:
            fh.write(e + nl)


@pytest.mark.parametrize("space", ["host", "device"])
@pytest.mark.parametrize("fmt_kind,dtype", [("csr", "float64"), ("csc", "float64"), ("dcsr", "float64"), ("csr", "float32")])
def test_read_mtx_general(tmp_path, fmt_kind, dtype, space):
    rng = np.random.default_rng(3)
    n, m = 300, 257
    entries = []
    for q in range(4000):                       # every numeral shape, duplicates included, ragged spacing, blank lines
        r, c = int(rng.integers(1, n + 1)), int(rng.integers(1, m + 1))
        num = NUMERALS[q % len(NUMERALS)] if q % 3 == 0 else repr(float(rng.normal() * 10 ** int(rng.integers(-8, 9))))
        entries.append(f"{r}  {c}\t{num}" + ("   " if q % 7 == 0 else ""))
    path = tmp_path / "a.mtx"
    _write_mtx(path, [n, m], entries, crlf=(fmt_kind == "dcsr"))
    fmt = {"csr": tb.CSR, "csc": tb.Format([tb.dense, tb.compressed], [1, 0]), "dcsr": tb.DCSR}[fmt_kind]
    tb.set_result_space(space)
    try:
        t = tb.read(path, fmt, np.dtype(dtype))
        dims, coords, vals = oracle.read_mtx(path)
        _check(t, fmt_kind, dims, coords, vals, np.dtype(dtype))
    finally:
        tb.set_result_space("host")


def test_read_mtx_symmetric_and_announced_count(tmp_path):
    rng = np.random.default_rng(5)
    n = 120
    seen, entries = set(), []
    while len(entries) < 700:
        r, c = int(rng.integers(1, n + 1)), int(rng.integers(1, n + 1))
        if r < c or (r, c) in seen:
            continue
        seen.add((r, c))
        entries.append(f"{r} {c} {float(rng.integers(-50, 50)) / 8}")
    path = tmp_path / "s.mtx"
    _write_mtx(path, [n, n], entries, symmetric=True)
    with open(path, "a") as fh:                 # lines past the announced count are ignored, as the reference ignores them
        fh.write("1 1 99.5\n\n")
    t = tb.read(path, tb.CSR)
    dims, coords, vals = oracle.read_mtx(path)
    _check(t, "csr", dims, coords, vals, np.float64)
    assert int(t.ct.vals_size) == 2 * len(entries) - sum(1 for e in entries if e.split()[0] == e.split()[1])


def test_read_tns_order3_and_rua32_shaped_file(tmp_path):
    rng = np.random.default_rng(9)
    path = tmp_path / "t.tns"
    with open(path, "w") as fh:
        for q in range(5000):
            i, j, k = (int(rng.integers(1, d + 1)) for d in (37, 50, 41))
            fh.write(f"{i} {j}  {k} {NUMERALS[q % len(NUMERALS)] if q % 4 == 0 else float(rng.integers(1, 1000)) / 16}\n")
    t = tb.read(path, tb.CSF3)
    dims, coords, vals = oracle.read_tns(path)
    _check(t, "csf3", dims, coords, vals, np.float64)
    # the reference's storage known answer for test/data/rua_32.mtx (tests-api.cpp:261-300): a file with the same entries, in
    # the file's column-major order, must read + pack to exactly that CSR
    pos, crd, vals = H.rua32_csr()
    rows = np.repeat(np.arange(32), np.diff(pos))
    order = np.lexsort((rows, crd))
    path = tmp_path / "rua_32.mtx"
    _write_mtx(path, [32, 32], [f"{rows[e] + 1} {crd[e] + 1} {vals[e]:.13e}" for e in order])
    t = tb.read(path, tb.CSR)
    p, c = t.level(1)
    assert np.array_equal(G.to_host(p), pos) and np.array_equal(G.to_host(c), crd) and np.array_equal(G.to_host(t.vals()), vals)
    # ... and it feeds the compute entry points directly
    x = np.arange(1, 33, dtype=np.float64)
    y = G.run("spmv", dict(dims=[32, 32], A_pos=G.to_host(p), A_crd=G.to_host(c), A_vals=G.to_host(t.vals()), x=x))
    assert np.array_equal(y, oracle.spmv(pos, crd, vals, x))


def test_read_refuses_what_the_reference_refuses(tmp_path):
    bad = tmp_path / "c.mtx"
    bad.write_text("%%MatrixMarket matrix coordinate complex general\n2 2 1\n1 1 1.0 0.0\n")
    with pytest.raises(tb.TacoError):
        tb.read(bad, tb.CSR)
    short = tmp_path / "short.mtx"
    short.write_text("%%MatrixMarket matrix coordinate real general\n3 3 5\n1 1 1.0\n")
    with pytest.raises(tb.TacoError):
        tb.read(short, tb.CSR)
    zero = tmp_path / "z.tns"
    zero.write_text("0 1 2.0\n")
    with pytest.raises(tb.TacoError):
        tb.read(zero, tb.CSR)
    with pytest.raises(tb.TacoError):
        tb.read(tmp_path / "missing.mtx", tb.CSR)
