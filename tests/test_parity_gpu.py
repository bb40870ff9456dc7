"""GPU parity tests (`-m gpu`): the CUDA path, called through the taco_tensor_t C ABI, against
  (a) the reference's own known-answer vectors,
  (b) outputs of the reference itself (tests/golden/*.npz),
  (c) the CPU oracle on seeded synthetic inputs (host-buffer AND device-resident calling conventions),
  (d) edge cases (empty tensors, empty rows, ragged widths, hub rows, rows with thousands of products).
Bar (north star): pos/crd bit-exact; values <= 1e-12 relative (fp64) / 1e-5 (fp32) -- and bit-exact wherever the
kernel keeps the reference's operation order (SpMV, SpMM non-hub rows, MTTKRP, TTV, TTM, SpAdd, SpGEMM).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle  # noqa: E402
import helpers as H  # noqa: E402
import gpu_util as G  # noqa: E402
import taco_b200 as tb  # noqa: E402
import synth  # noqa: E402  (tests/synth.py: workload generators)
from taco_b200 import formats  # noqa: E402

pytestmark = pytest.mark.gpu

SPACES = ["host", "device"]


def place(w, space):
    return G.to_device(w) if space == "device" else w


# ---------------------------------------------------------------------------------------------------------
# (a) reference known-answer vectors through the C ABI
# ---------------------------------------------------------------------------------------------------------
def test_kat_spmv():
    p, c, v = formats.csr_from_dense(H.d33a())
    y = G.run("spmv", dict(dims=[3, 3], A_pos=p, A_crd=c, A_vals=v, x=H.d3b()))
    assert y.tolist() == [0, 0, 18]


def test_kat_matrix_add_structure():
    ap, ac, av = formats.csr_from_dense(H.d33a())
    bp, bc, bv = formats.csr_from_dense(H.d33b())
    pos, crd, vals = G.run("spadd", dict(dims=[3, 3], A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv))
    assert pos.tolist() == [0, 2, 2, 5] and crd.tolist() == [0, 1, 0, 1, 2] and vals.tolist() == [10, 22, 3, 30, 4]
    ap, ac, av = formats.csr_from_dense(H.d34a())
    bp, bc, bv = formats.csr_from_dense(H.d34b())
    pos, crd, vals = G.run("spadd", dict(dims=[3, 4], A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv),
                           phases="separate")
    assert pos.tolist() == [0, 3, 3, 6] and crd.tolist() == [0, 2, 3, 0, 2, 3] and vals.tolist() == [4, 3, 3, 8, 5, 5]


def test_kat_matrix_mul():
    ap, ac, av = formats.csr_from_dense(H.d33a())
    C = G.run("spmm", dict(dims=[3, 3, 3], A_pos=ap, A_crd=ac, A_vals=av, B=H.d33b().reshape(-1)))
    assert C.tolist() == [0, 0, 0, 0, 0, 0, 30, 180, 0]
    Ct = G.run("spmm", dict(dims=[3, 3, 3], A_pos=ap, A_crd=ac, A_vals=av, B=H.d33b().reshape(-1)), colmajor_c=True)
    assert Ct.reshape(3, 3).T.reshape(-1).tolist() == C.tolist()
    bp, bc, bv = formats.csr_from_dense(H.d33b())
    pos, crd, vals = G.run("spgemm", dict(dims=[3, 3, 3], A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv))
    assert pos.tolist() == [0, 0, 0, 2] and crd.tolist() == [0, 1] and vals.tolist() == [30, 180]


def test_kat_mttkrp_ttv_ttm():
    t = formats.coo_to_csf3(*H.d233a())
    A = G.run("mttkrp", dict(dims=[2, 3, 3, 3], C=H.d33a().reshape(-1), D=H.d33b().reshape(-1), **t))
    assert A.tolist() == [0, 80, 0, 180, 0, 0]
    A = G.run("ttm", dict(dims=[2, 3, 3, 3], C=H.d33a().reshape(-1), **t))
    assert A.tolist() == [0, 4, 0, 0, 0, 0, 12, 0, 16, 0, 0, 0, 0, 0, 0, 21, 12, 28]
    t3 = formats.coo_to_csf3(*H.d333a())
    A = G.run("ttv", dict(dims=[3, 3, 3], c=H.d3b(), **t3))
    assert A.tolist() == [4, 0, 12, 0, 0, 33, 0, 24, 0]
    # tests-parafac.cpp:157-187 (mttkrp2 / mttkrp3: mode-J and mode-K MTTKRP = the standard kernel over the permuted storage)
    a, b, c, v = H.d333a()
    A = G.run("mttkrp", dict(dims=[3, 3, 3, 3], C=H.d33a().reshape(-1), D=H.d33b().reshape(-1), **formats.coo_to_csf3(b, a, c, v)))
    assert A.tolist() == [0, 80, 0, 0, 0, 0, 0, 240, 0]
    A = G.run("mttkrp", dict(dims=[3, 3, 3, 3], C=H.d33a().reshape(-1), D=H.d33b().reshape(-1), **formats.coo_to_csf3(c, a, b, v)))
    assert A.tolist() == [0, 80, 0, 0, 120, 0, 0, 240, 0]


# ---------------------------------------------------------------------------------------------------------
# (b) outputs of the reference itself
# ---------------------------------------------------------------------------------------------------------
def _inputs(g):
    return {k: v for k, v in g.items() if not k.startswith("out_")}


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("spmv"))
def test_golden_spmv(name, space):
    g = H.load_golden(name)
    y = G.run("spmv", place(_inputs(g), space))
    if name.endswith("_default") or "_int_" in name:
        assert np.array_equal(y, g["out_y"])          # same operation order as the reference's C kernel
    else:
        H.assert_close(y, g["out_y"], y.dtype)


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("spmm"))
def test_golden_spmm(name, space):
    g = H.load_golden(name)
    C = G.run("spmm", place(_inputs(g), space))
    assert np.array_equal(C, g["out_C"])
    n, m, K = g["dims"]
    Ct = G.run("spmm", place(_inputs(g), space), colmajor_c=True)
    assert np.array_equal(Ct.reshape(K, n).T.reshape(-1), g["out_C"])


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("dcsr_spmm"))
def test_golden_spmm_dcsr(name, space):
    # SpMM with a doubly compressed operand (the reference's spmmDCSRGPU statement): same kernel as CSR after the
    # level-0 row list has been expanded on the device -> bit-identical to the reference, absent rows are zero rows
    g = H.load_golden(name)
    C = G.run("spmm_dcsr", place(_inputs(g), space))
    assert np.array_equal(C, g["out_C"])
    n, m, K = g["dims"]
    Ct = G.run("spmm_dcsr", place(_inputs(g), space), colmajor_c=True)
    assert np.array_equal(Ct.reshape(K, n).T.reshape(-1), g["out_C"])


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("keep", [0.0, 0.02, 0.6, 1.0])
def test_oracle_spmm_dcsr(space, keep):
    # R-MAT rows (hub rows included) with a fraction `keep` of the rows stored; keep = 0 is the empty operand
    w = synth.make("spmm", None, scale=12, K=64, dtype="float32")
    n, m, K = w["dims"]
    rng = np.random.default_rng(11)
    lens = np.diff(w["A_pos"])
    stored = np.flatnonzero((rng.random(n) < keep) & (lens > 0)).astype(np.int32)
    take = np.concatenate([np.arange(w["A_pos"][r], w["A_pos"][r + 1]) for r in stored]) if stored.size else np.zeros(0, np.int64)
    pos2 = np.zeros(stored.size + 1, np.int32)
    np.cumsum(lens[stored], out=pos2[1:])
    d = dict(dims=w["dims"], A1_pos=np.array([0, stored.size], np.int32), A1_crd=stored, A2_pos=pos2,
             A2_crd=w["A_crd"][take].astype(np.int32), A_vals=w["A_vals"][take], B=w["B"])
    C = G.run("spmm_dcsr", place(d, space)).reshape(n, K)
    want = oracle.spmm_dcsr(n, d["A1_pos"], d["A1_crd"], d["A2_pos"], d["A2_crd"], d["A_vals"], w["B"].reshape(m, K))
    short = np.ones(n, bool)
    short[stored] = lens[stored] <= 128
    assert np.array_equal(C[short], want[short])
    H.assert_close(C.reshape(-1), want.reshape(-1), np.float32)


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("densesddmm"))
def test_golden_sddmm_dense(name, space):
    # the reference's sddmmGPU statement: dense result, D indexed (contraction, column)
    g = H.load_golden(name)
    A = G.run("sddmm_dense", place(_inputs(g), space))
    if "_int_" in name:
        assert np.array_equal(A, g["out_A"])
    else:
        H.assert_close(A, g["out_A"], g["B_vals"].dtype, scale=float(np.abs(g["out_A"]).max()))


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("sddmm"))
def test_golden_sddmm(name, space):
    g = H.load_golden(name)
    pos, crd, vals = G.run("sddmm", place(_inputs(g), space))
    assert np.array_equal(pos, g["out_A_pos"]) and np.array_equal(crd, g["out_A_crd"])
    if "_int_" in name:
        assert np.array_equal(vals, g["out_A_vals"])
    else:   # the K-contraction is a shuffle tree here: reduction reordering, north-star tolerance
        H.assert_close(vals, g["out_A_vals"], vals.dtype, scale=np.abs(g["out_A_vals"]).max())


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("fam", ["mttkrp", "ttv", "ttm"])
@pytest.mark.parametrize("tag", ["int", "frac"])
def test_golden_csf(fam, tag, space):
    g = H.load_golden(f"{fam}_{tag}_f64")
    A = G.run(fam, place(_inputs(g), space))
    assert np.array_equal(A, g["out_A"])


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("phases", ["evaluate", "separate"])
@pytest.mark.parametrize("name", H.golden_cases("spadd") + H.golden_cases("spgemm"))
def test_golden_sparse_output(name, space, phases):
    g = H.load_golden(name)
    fam = name.split("_")[0]
    pos, crd, vals = G.run(fam, place(_inputs(g), space), phases=phases)
    assert np.array_equal(pos, g["out_C_pos"]), "pos must be bit-exact"
    assert np.array_equal(crd, g["out_C_crd"]), "crd must be bit-exact"
    assert np.array_equal(vals, g["out_C_vals"])


# ---------------------------------------------------------------------------------------------------------
# (c) oracle on seeded synthetic inputs
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_oracle_spmv(space, dtype):
    w = synth.make("spmv", None, n=200_003, deg=10, dtype=dtype)
    y = G.run("spmv", place(w, space))
    assert np.array_equal(y, oracle.spmv(w["A_pos"], w["A_crd"], w["A_vals"], w["x"]))


def test_oracle_spmv_powerlaw_long_rows():
    # R-MAT structure: empty rows, rows far longer than one staging tile (slow path with carries)
    pos, crd, vals = synth.csr_rmat(synth.backend(None), 16, 16, 123, np.float64)
    assert np.diff(pos).max() > 2048
    x = synth.dense(synth.backend(None), 1 << 16, 1, 5, np.float64)
    w = dict(dims=[1 << 16, 1 << 16], A_pos=pos, A_crd=crd, A_vals=vals, x=x)
    for space in SPACES:
        y = G.run("spmv", place(w, space))
        assert np.array_equal(y, oracle.spmv(pos, crd, vals, x))


def _spmm_check(w, C, dtype):
    n, m, K = w["dims"]
    want = oracle.spmm(w["A_pos"], w["A_crd"], w["A_vals"], w["B"].reshape(m, K))
    C = C.reshape(n, K)
    deg = np.diff(w["A_pos"])
    short = deg <= 128
    assert np.array_equal(C[short], want[short]), "non-hub rows keep the reference's operation order: bit-exact"
    if (~short).any():   # hub rows are split across slots and combined with red.global.add: reordered sums
        H.assert_close(C[~short], want[~short], dtype)
    return int((~short).sum())


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("K,dtype", [(128, "float32"), (64, "float64"), (100, "float32"), (33, "float32"), (7, "float64")])
def test_oracle_spmm(space, K, dtype):
    w = synth.make("spmm", None, scale=15, K=K, dtype=dtype)
    C = G.run("spmm", place(w, space))
    hubs = _spmm_check(w, C, dtype)
    assert hubs > 0, "the R-MAT case is meant to exercise the hub-row path"


@pytest.mark.parametrize("K,dtype", [(128, "float32"), (40, "float64")])
def test_spmm_long_rows_deterministic(K, dtype):
    # long (hub) rows go through the column-panel schedule: partial sums per work item, combined in position order by
    # one warp per row -- no atomics, so repeated runs must agree bit for bit (and with the column-major result)
    w = G.to_device(synth.make("spmm", None, scale=14, K=K, dtype=dtype))
    assert int((w["A_pos"][1:] - w["A_pos"][:-1]).max()) > 512
    first = G.run("spmm", w)
    for _ in range(4):
        assert np.array_equal(G.run("spmm", w), first)
    n = int(w["dims"][0])
    Ct = G.run("spmm", w, colmajor_c=True)
    assert np.array_equal(Ct.reshape(K, n).T.reshape(-1), first)


@pytest.mark.parametrize("K,dtype", [(128, "float32"), (48, "float64")])
def test_oracle_spmm_host_pipeline(K, dtype, monkeypatch):
    # host operands through the chunked upload / compute / download pipeline (forced for this small case); pinned and
    # pageable buffers, plus B registered resident
    monkeypatch.setenv("TACO_B200_PIPELINE_MIN_BYTES", "0")
    w = synth.make("spmm", None, scale=13, K=K, dtype=dtype)
    C = G.run("spmm", w)
    _spmm_check(w, C, dtype)
    wp = {}
    for k, v in w.items():
        if k == "dims":
            wp[k] = v
        else:
            a = tb.pinned_empty(v.shape, v.dtype)
            a[...] = v
            wp[k] = a
    C = G.run("spmm", wp)
    _spmm_check(w, C, dtype)
    from taco_b200 import _lib
    _lib.check(_lib.lib.taco_b200_make_resident(wp["B"].ctypes.data, wp["B"].nbytes))
    try:
        C = G.run("spmm", wp)
        _spmm_check(w, C, dtype)
    finally:
        _lib.lib.taco_b200_invalidate(wp["B"].ctypes.data)
        for k, v in wp.items():
            if k != "dims":
                tb.pinned_free(v)


@pytest.mark.parametrize("K,dtype", [(64, "float32"), (64, "float64"), (20, "float32"), (5, "float64"), (256, "float32")])
def test_oracle_sddmm(K, dtype):
    w = synth.make("sddmm", None, n=30_011, deg=20, K=K, dtype=dtype)
    n = w["dims"][0]
    ap, ac, av = oracle.sddmm(w["B_pos"], w["B_crd"], w["B_vals"], w["C"].reshape(n, K), w["D"].reshape(n, K))
    for space in SPACES:
        pos, crd, vals = G.run("sddmm", place(w, space))
        assert np.array_equal(pos, ap) and np.array_equal(crd, ac)
        H.assert_close(vals, av, dtype)


@pytest.mark.parametrize("R,dtype", [(32, "float64"), (16, "float64"), (40, "float32")])
def test_oracle_mttkrp(R, dtype):
    w = synth.make("mttkrp", None, I=20_000, K=3_000, L=2_500, nnz=400_000, R=R, dtype=dtype)
    I, K, L, _ = w["dims"]
    want = oracle.mttkrp(w, w["C"].reshape(K, R), w["D"].reshape(L, R), I)
    for space in SPACES:
        A = G.run("mttkrp", place(w, space))
        assert np.array_equal(A.reshape(I, R), want)


def test_oracle_mttkrp_long_fibers_and_gaps():
    # few slices (gaps in mode 0), long fibers: exercises the >32-leaf path and the zero fill of unoccupied rows
    w = synth.make("mttkrp", None, I=5_000, K=40, L=3_000, nnz=150_000, R=32, dtype="float64")
    keep = w["B1_crd"] % 3 != 1
    i, k, l, v = formats.csf3_to_coo(w)
    sel = keep[np.searchsorted(w["B1_crd"], i)]
    t = formats.coo_to_csf3(i[sel], k[sel], l[sel], v[sel])
    w2 = dict(dims=w["dims"], C=w["C"], D=w["D"], **t)
    want = oracle.mttkrp(t, w["C"].reshape(40, 32), w["D"].reshape(3000, 32), 5000)
    A = G.run("mttkrp", G.to_device(w2))
    assert np.array_equal(A.reshape(5000, 32), want)


@pytest.mark.parametrize("dtype,R", [("float64", 32), ("float32", 40)])
def test_mttkrp_hub_slices_deterministic(dtype, R):
    # slices of thousands of leaves are split across 64-leaf slots; the pieces are added into the row in slot order
    # (owner stores, every later piece waits for its predecessor): no atomics on values, so repeated runs agree bit for
    # bit; whole slices keep the reference's order exactly.  Long fibers: C(k,:) is gathered once per fiber.
    I, K, L = 9, 7, 5000
    w = synth.make("mttkrp", None, I=I, K=K, L=L, nnz=150_000, R=R, dtype=dtype)
    t = {k: v for k, v in w.items() if k.startswith("B")}
    assert np.diff(t["B3_pos"][t["B2_pos"]]).max() > 2000 and np.diff(t["B3_pos"]).mean() > 100
    wd = G.to_device(dict(dims=w["dims"], C=w["C"], D=w["D"], **t))
    first = G.run("mttkrp", wd)
    for _ in range(4):
        assert np.array_equal(G.run("mttkrp", wd), first)
    want = oracle.mttkrp(t, w["C"].reshape(K, R), w["D"].reshape(L, R), I)
    H.assert_close(first, want.reshape(-1), np.dtype(dtype))


@pytest.mark.parametrize("dtype,R", [("float64", 32), ("float32", 16), ("float64", 5)])
def test_oracle_mttkrp_host_pipeline(dtype, R, monkeypatch):
    # host operands through the slice-chunked upload / rebase / kernel / download pipeline (forced for this small case):
    # empty slices and rows in between, a hub slice, pageable and pinned buffers; bit-identical to the unpipelined call
    monkeypatch.setenv("TACO_B200_PIPELINE_MIN_BYTES", "0")
    w = synth.make("mttkrp", None, I=5_000, K=40, L=3_000, nnz=400_000, R=R, dtype=dtype)
    keep = np.ones(w["B1_crd"].shape[0], bool)
    keep[:3] = False                                       # leading rows of A without a slice
    keep[100:400] = False                                  # a gap
    keep[-5:] = False                                      # trailing rows without a slice
    i, k, l, v = formats.csf3_to_coo(w)
    sel = keep[np.searchsorted(w["B1_crd"], i)]
    hub = np.arange(2500, dtype=np.int64)                  # one slice with 2500 leaves (> 512: split across slots)
    i2 = np.concatenate([i[sel], np.full(hub.size, 777)])
    k2 = np.concatenate([k[sel], hub % 40])
    l2 = np.concatenate([l[sel], hub % 3000])
    v2 = np.concatenate([v[sel], np.ones(hub.size, v.dtype)])
    flat = (i2.astype(np.int64) * 40 + k2) * 3000 + l2
    _, first = np.unique(flat, return_index=True)
    t = formats.coo_to_csf3(i2[first], k2[first], l2[first], v2[first])
    want = oracle.mttkrp(t, w["C"].reshape(40, R), w["D"].reshape(3000, R), 5000)
    w2 = dict(dims=w["dims"], C=w["C"], D=w["D"], **t)
    A = G.run("mttkrp", w2).reshape(5000, R)
    short = np.ones(5000, bool)
    short[777] = False
    assert np.array_equal(A[short], want[short])
    H.assert_close(A.reshape(-1), want.reshape(-1), np.dtype(dtype))
    wp = {}
    for key, val in w2.items():
        if key == "dims":
            wp[key] = val
        else:
            a = tb.pinned_empty(val.shape, val.dtype)
            a[...] = val
            wp[key] = a
    try:
        A2 = G.run("mttkrp", wp).reshape(5000, R)
        assert np.array_equal(A2[short], want[short])
        H.assert_close(A2.reshape(-1), want.reshape(-1), np.dtype(dtype))
    finally:
        for key, val in wp.items():
            if key != "dims":
                tb.pinned_free(val)
    monkeypatch.setenv("TACO_B200_PIPELINE_MIN_BYTES", str(1 << 40))          # the unpipelined path on the same operands
    A3 = G.run("mttkrp", w2).reshape(5000, R)
    assert np.array_equal(A3[short], A[short])


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("mode", [1, 2])
def test_mttkrp_mode_j_and_k_over_permuted_storage(space, mode):
    # the other two MTTKRPs of a CP-ALS sweep (tests-parafac.cpp:157-187): B stored with the result's mode first
    rng = np.random.default_rng(40 + mode)
    dims, R = (37, 23, 29), 8
    n = 4000
    c = [rng.integers(0, d, n) for d in dims]
    flat = np.ravel_multi_index(c, dims)
    _, first = np.unique(flat, return_index=True)
    c = [x[first] for x in c]
    v = np.floor(rng.random(first.size) * 5 + 1)
    ordering = [1, 0, 2] if mode == 1 else [2, 0, 1]
    t = formats.coo_to_csf3(c[ordering[0]], c[ordering[1]], c[ordering[2]], v)           # CSF of the permuted tensor
    Cm = np.floor(rng.random((dims[ordering[1]], R)) * 4)
    Dm = np.floor(rng.random((dims[ordering[2]], R)) * 4)
    dense = np.zeros(dims)
    dense[tuple(c)] = v
    want = np.einsum("kil,kr,lr->ir" if mode == 1 else "kli,kr,lr->ir", dense, Cm, Dm)
    arrs = place(dict(C=Cm.reshape(-1), D=Dm.reshape(-1), **t), space)
    tb.set_result_space("device" if space == "device" else "host")
    try:
        B = tb.makeCSF3("B", list(dims), arrs, ordering)
        Ct = tb.makeDense("C", [dims[ordering[1]], R], arrs["C"])
        Dt = tb.makeDense("D", [dims[ordering[2]], R], arrs["D"])
        A = tb.Tensor("A", [dims[mode], R], tb.Format([tb.dense, tb.dense]), np.float64)
        expr = "A(i,j) = B(k,i,l) * C(k,j) * D(l,j)" if mode == 1 else "A(i,j) = B(k,l,i) * C(k,j) * D(l,j)"
        tb.compile(expr, A, B, Ct, Dt)(A, B, Ct, Dt)
        if space == "device":
            tb.synchronize()
        got = G.to_host(A.vals()).reshape(dims[mode], R)
    finally:
        tb.set_result_space("host")
    assert np.array_equal(got, want)


def test_oracle_ttv_ttm():
    w = synth.make("mttkrp", None, I=3_000, K=500, L=800, nnz=200_000, R=16, dtype="float64")
    I, K, L, R = w["dims"]
    t = {k: v for k, v in w.items() if k.startswith("B")}
    c = w["D"][:L].copy()
    A = G.run("ttv", G.to_device(dict(dims=[I, K, L], c=c, **t)))
    assert np.array_equal(A.reshape(I, K), oracle.ttv(t, c, I, K))
    A = G.run("ttm", G.to_device(dict(dims=[I, K, L, R], C=w["D"], **t)))
    assert np.array_equal(A.reshape(I, K, R), oracle.ttm(t, w["D"].reshape(L, R), I, K))


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("dtype,R", [("float64", 32), ("float64", 7), ("float32", 64), ("float32", 5)])
def test_oracle_ttv_ttm_long_fibers(space, dtype, R):
    # few (i,j) cells, long fibers: fibers of thousands of leaves span several 2048-leaf SpMV tiles (hand-over between
    # tiles) and several 64-leaf SpMM slots (hub rows, combined with red.global.add); empty slices and cells in between
    I, K, L = 37, 23, 6000
    w = synth.make("mttkrp", None, I=I, K=K, L=L, nnz=600_000, R=R, dtype=dtype)
    t = {k: v for k, v in w.items() if k.startswith("B")}
    lens = np.diff(t["B3_pos"])
    assert lens.max() > 600
    c = w["D"].reshape(L, R)[:, 0].copy()
    A = G.run("ttv", place(dict(dims=[I, K, L], c=c, **t), space)).reshape(I, K)
    want = oracle.ttv(t, c, I, K)
    H.assert_close(A.reshape(-1), want.reshape(-1), np.dtype(dtype))
    A = G.run("ttm", place(dict(dims=[I, K, L, R], C=w["D"], **t), space)).reshape(I, K, R)
    want = oracle.ttm(t, w["D"].reshape(L, R), I, K)
    H.assert_close(A.reshape(-1), want.reshape(-1), np.dtype(dtype))
    # a tensor whose fibers are all short keeps the reference's order exactly
    w = synth.make("mttkrp", None, I=400, K=300, L=L, nnz=300_000, R=R, dtype=dtype)
    t = {k: v for k, v in w.items() if k.startswith("B")}
    A = G.run("ttm", place(dict(dims=[400, 300, L, R], C=w["D"], **t), space)).reshape(400, 300, R)
    assert np.array_equal(A, oracle.ttm(t, w["D"].reshape(L, R), 400, 300))
    c = w["D"].reshape(L, R)[:, 0].copy()
    A = G.run("ttv", place(dict(dims=[400, 300, L], c=c, **t), space)).reshape(400, 300)
    assert np.array_equal(A, oracle.ttv(t, c, 400, 300))


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_oracle_spadd(space, dtype):
    w = synth.make("spadd", None, n=100_003, deg=10, dtype=dtype)
    w["B_vals"][::7] = -1.0                       # some sums cancel to an explicit zero when columns coincide
    cp, cc, cv = oracle.spadd(w["A_pos"], w["A_crd"], w["A_vals"], w["B_pos"], w["B_crd"], w["B_vals"])
    pos, crd, vals = G.run("spadd", place(w, space), phases="separate")
    assert np.array_equal(pos, cp) and np.array_equal(crd, cc) and np.array_equal(vals, cv)


@pytest.mark.parametrize("phases", ["evaluate", "separate"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_oracle_spadd_evaluate_and_skewed_rows(phases, dtype):
    # evaluate = the one-pass union kernel (tickets + look-back); R-MAT operands put hub rows (> staging capacity) next to
    # empty ones, so staged and direct row blocks alternate inside one launch; many columns coincide
    xp = synth.backend(None)
    ap, ac, av = synth.csr_rmat(xp, 13, 12, 91, np.dtype(dtype))
    bp, bc, bv = synth.csr_rmat(xp, 13, 20, 91, np.dtype(dtype))       # same seed, more edges: heavy overlap with A
    n = 1 << 13
    assert np.diff(ap).max() > 1024 and (np.diff(ap) == 0).sum() > 100
    cp, cc, cv = oracle.spadd(ap, ac, av, bp, bc, bv)
    assert len(cc) < len(ac) + len(bc)
    w = dict(dims=[n, n], A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv)
    for space in SPACES:
        pos, crd, vals = G.run("spadd", place(w, space), phases=phases)
        assert np.array_equal(pos, cp) and np.array_equal(crd, cc) and np.array_equal(vals, cv)
    # uniform operands at a size that is not a multiple of the row block
    w = synth.make("spadd", None, n=77_777, deg=13, dtype=dtype)
    cp, cc, cv = oracle.spadd(w["A_pos"], w["A_crd"], w["A_vals"], w["B_pos"], w["B_crd"], w["B_vals"])
    pos, crd, vals = G.run("spadd", G.to_device(w), phases=phases)
    assert np.array_equal(pos, cp) and np.array_equal(crd, cc) and np.array_equal(vals, cv)


@pytest.mark.parametrize("space", SPACES)
def test_oracle_spgemm(space):
    w = synth.make("spgemm", None, n=40_009, deg=10, dtype="float64")
    n = w["dims"][0]
    cp, cc, cv = oracle.spgemm(w["A_pos"], w["A_crd"], w["A_vals"], w["B_pos"], w["B_crd"], w["B_vals"], n)
    pos, crd, vals = G.run("spgemm", place(w, space), phases="separate")
    assert np.array_equal(pos, cp) and np.array_equal(crd, cc), "structure must be bit-exact"
    assert np.array_equal(vals, cv)


def test_oracle_spgemm_powerlaw_all_bins():
    # R-MAT x R-MAT: rows with > 256 products (CTA-sort bin) and > 8192 products (bitmap bin), many collisions
    xp = synth.backend(None)
    ap, ac, av = synth.csr_rmat(xp, 12, 24, 77, np.float64)
    bp, bc, bv = synth.csr_rmat(xp, 12, 24, 78, np.float64)
    n = 1 << 12
    csum = np.concatenate([[0], np.cumsum(np.diff(bp)[ac])])
    ub = csum[ap[1:]] - csum[ap[:-1]]                                   # products per row
    assert ub.max() > 8192 and (ub > 256).sum() > 10
    cp, cc, cv = oracle.spgemm(ap, ac, av, bp, bc, bv, n)
    w = dict(dims=[n, n, n], A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv)
    for space in SPACES:
        pos, crd, vals = G.run("spgemm", place(w, space))
        assert np.array_equal(pos, cp) and np.array_equal(crd, cc)
        assert np.array_equal(vals, cv)


# ---------------------------------------------------------------------------------------------------------
# (d) edge cases
# ---------------------------------------------------------------------------------------------------------
def _empty_csr(n):
    return np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float64)


@pytest.mark.parametrize("space", SPACES)
def test_empty_operands(space):
    p, c, v = _empty_csr(5)
    y = G.run("spmv", place(dict(dims=[5, 4], A_pos=p, A_crd=c, A_vals=v, x=np.ones(4)), space))
    assert y.tolist() == [0] * 5
    C = G.run("spmm", place(dict(dims=[5, 4, 8], A_pos=p, A_crd=c, A_vals=v, B=np.ones(32)), space))
    assert C.tolist() == [0] * 40
    q, d, u = formats.csr_from_dense(np.array([[0, 1.0, 0, 2], [0, 0, 0, 0], [3, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]]))
    pos, crd, vals = G.run("spadd", place(dict(dims=[5, 4], A_pos=p, A_crd=c, A_vals=v, B_pos=q, B_crd=d, B_vals=u), space))
    assert pos.tolist() == q.tolist() and crd.tolist() == d.tolist() and vals.tolist() == u.tolist()
    pos, crd, vals = G.run("spadd", place(dict(dims=[5, 4], A_pos=p, A_crd=c, A_vals=v, B_pos=p, B_crd=c, B_vals=v), space))
    assert pos.tolist() == [0] * 6 and crd.size == 0 and vals.size == 0
    p4, c4, v4 = _empty_csr(4)
    pos, crd, vals = G.run("spgemm", place(dict(dims=[5, 4, 4], A_pos=q, A_crd=d, A_vals=u, B_pos=p4, B_crd=c4, B_vals=v4), space))
    assert pos.tolist() == [0] * 6 and crd.size == 0
    pos, crd, vals = G.run("sddmm", place(dict(dims=[5, 4, 3], B_pos=p, B_crd=c, B_vals=v, C=np.ones(15), D=np.ones(12)), space))
    assert pos.tolist() == [0] * 6 and vals.size == 0


def test_empty_csf():
    t = formats.coo_to_csf3(np.zeros(0, int), np.zeros(0, int), np.zeros(0, int), np.zeros(0))
    A = G.run("mttkrp", dict(dims=[4, 3, 3, 8], C=np.ones(24), D=np.ones(24), **t))
    assert A.tolist() == [0] * 32


def test_leading_trailing_empty_rows_and_single_row():
    d = np.zeros((70, 9))
    d[33] = np.arange(1, 10)
    d[34, 2] = 5
    p, c, v = formats.csr_from_dense(d)
    x = np.arange(9, dtype=np.float64) + 1
    y = G.run("spmv", dict(dims=[70, 9], A_pos=p, A_crd=c, A_vals=v, x=x))
    assert np.array_equal(y, d @ x)
    B = np.arange(9 * 8, dtype=np.float64).reshape(9, 8)
    C = G.run("spmm", dict(dims=[70, 9, 8], A_pos=p, A_crd=c, A_vals=v, B=B.reshape(-1)))
    assert np.array_equal(C.reshape(70, 8), d @ B)


def test_dimension_and_type_errors():
    p, c, v = formats.csr_from_dense(np.eye(3))
    A = tb.makeCSR("A", [3, 3], p, c, v)
    x = tb.makeDense("x", [4], np.ones(4))
    y = tb.Tensor("y", [3], tb.Format([tb.dense]))
    k = tb.compile("y(i) = A(i,j) * x(j)", y, A, tb.makeDense("x", [3], np.ones(3)))
    with pytest.raises(tb.TacoError) as ei:
        k(y, A, x)
    assert ei.value.code == 3
    with pytest.raises(tb.TacoError):   # compute without assemble: result has no storage
        k.compute(tb.Tensor("y", [3], tb.Format([tb.dense])), A, tb.makeDense("x", [3], np.ones(3)))


def test_resident_cache_and_launch_count():
    from taco_b200 import _lib
    import ctypes
    w = synth.make("spmv", None, n=50_000, deg=10)
    for key in ("A_pos", "A_crd", "A_vals"):
        a = w[key]
        tb._lib.check(_lib.lib.taco_b200_make_resident(ctypes.c_void_p(a.ctypes.data), a.nbytes))
    before = tb.launch_count()
    y1 = G.run("spmv", w)
    assert tb.launch_count() > before, "the library must have launched its own kernels"
    w["A_vals"][:] = 0                      # host copy changes, resident mirror does not ...
    y2 = G.run("spmv", w)
    assert np.array_equal(y1, y2)
    _lib.lib.taco_b200_invalidate(ctypes.c_void_p(w["A_vals"].ctypes.data))   # ... until it is invalidated
    y3 = G.run("spmv", w)
    assert not y3.any()
    _lib.lib.taco_b200_drop_all_resident()


def test_partition_pos_device_matches_host():
    import torch
    w = synth.make("spmm", None, scale=14, K=4)
    n = w["dims"][0]
    host = tb.partition_pos(w["A_pos"], n, 8)
    dev = tb.partition_pos(torch.as_tensor(w["A_pos"]).cuda(), n, 8)
    assert host.tolist() == dev.tolist()
    sizes = np.diff(w["A_pos"][host])
    assert sizes.max() - sizes.min() <= np.diff(w["A_pos"]).max()      # nnz-balanced up to one row


# ---------------------------------------------------------------------------------------------------------
# (e) drop-in: an ordinary taco C++ program on the UNMODIFIED reference library, kernels forwarded to libtaco_b200
# ---------------------------------------------------------------------------------------------------------
def test_dropin_demo_through_unmodified_reference():
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    exe = os.path.join(root, "oracle", "_ref", "taco_dropin_demo")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref (the compiled reference) was not shipped with this snapshot")
    r = subprocess.run([exe, tb.LIB_PATH], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("expr,fmts,files", [
    ("y(i) = A(i,j) * x(j)", dict(y="d", A="ds", x="d"), dict(A=(300, 200), x=(200,))),
    ("C(i,k) = A(i,j) * B(j,k)", dict(C="dd", A="ds", B="dd"), dict(A=(120, 90), B=(90, 16))),
    ("C(i,j) = A(i,j) + B(i,j)", dict(C="ds", A="ds", B="ds"), dict(A=(150, 130), B=(150, 130))),
    ("C(i,k) = A(i,j) * B(j,k)", dict(C="ds", A="ds", B="ds"), dict(A=(80, 70), B=(70, 60))),
])
def test_cli_dropin_read_source(tmp_path, expr, fmts, files):
    """The reference's own `taco` command-line tool (tools/taco.cpp, compiled unmodified into oracle/_ref) evaluates the
    statement twice -- with its C codegen and with `-read-source=<stub>` -- and `-verify` compares the two results
    (tools/taco.cpp:1212-1259).  The stub forwards assemble / compute to libtaco_b200, so the CLI is a drop-in."""
    import subprocess
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    exe = os.path.join(root, "oracle", "_ref", "taco")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/taco (the compiled reference CLI) was not shipped with this snapshot")
    sys.path.insert(0, os.path.join(root, "tools"))
    import emit_stub
    rng = np.random.default_rng(len(expr))
    args = [exe, expr]
    for name, f in fmts.items():
        args.append(f"-f={name}:{f}")
    for name, shape in files.items():
        path = tmp_path / f"{name}.tns"                       # FROSTT coordinates, 1-based
        dense_operand = "s" not in fmts[name]
        with open(path, "w") as fh:
            for idx in np.ndindex(*shape):
                last = all(i == n - 1 for i, n in zip(idx, shape))      # pins the dimensions the CLI infers from the file
                if dense_operand or last or rng.random() < 0.15:
                    fh.write(" ".join(str(i + 1) for i in idx) + f" {int(rng.integers(1, 9))}\n")
        args.append(f"-i={name}:{path}")
    stub = tmp_path / "stub.c"
    stub.write_text(emit_stub.stub_source(expr, ",".join(f"{n}:{f}" for n, f in fmts.items()), "f64"))
    args += [f"-read-source={stub}", "-verify"]
    before = tb.launch_count()
    env = dict(os.environ, TACO_B200_LIB=tb.LIB_PATH, TACO_CFLAGS="-O3 -std=gnu99", TMPDIR=str(tmp_path))
    r = subprocess.run(args, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "Verifying... done" in r.stdout and "differ" not in r.stderr, r.stdout + r.stderr


# ---------------------------------------------------------------------------------------------------------
# blocked SpMV / SpMM (BCSR = {Dense,Compressed,Dense,Dense}), SURVEY.md 8(f) item 1
# ---------------------------------------------------------------------------------------------------------
def _uses_tensor_cores(br, bc, K, dtype):
    return np.dtype(dtype) == np.float32 and (br, bc) in ((16, 16), (32, 32)) and K % 4 == 0


def _bspmm_check(C, want, br, bc, K, dtype):
    if _uses_tensor_cores(br, bc, K, dtype):
        # tcgen05 path: three TF32 products per fp32 product, fp32 accumulation, reduction order differs from the
        # reference -> north-star fp32 tolerance (1e-5), measured against the largest entry of the result row block
        H.assert_close(C.reshape(-1), want.reshape(-1), np.float32, scale=float(np.max(np.abs(want))) if want.size else 1.0)
    else:
        assert np.array_equal(C.reshape(-1), want.reshape(-1)), "CUDA-core blocked SpMM keeps the reference's order"


def test_kat_bspmv():
    # tests-expr_storage.cpp:939-960
    pos, crd, blocks = H.d3322a()
    a = G.run("bspmv", dict(dims=[3, 3, 2, 2], A_pos=pos, A_crd=crd, A_vals=blocks.reshape(-1), c=H.d32b().reshape(-1)))
    H.assert_close(a, np.array([88.2, 96.4, 0.0, 0.0, 319.4, 335.8]), np.float64)
    assert np.array_equal(a, oracle.bspmv(pos, crd, blocks, H.d32b(), 2, 2).reshape(-1))


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("bspmv"))
def test_golden_bspmv(name, space):
    g = H.load_golden(name)
    a = G.run("bspmv", place(_inputs(g), space))
    assert np.array_equal(a, g["out_a"])


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("bspmm"))
def test_golden_bspmm(name, space):
    g = H.load_golden(name)
    Mb, Nb, br, bc, K = [int(x) for x in g["dims"]]
    C = G.run("bspmm", place(_inputs(g), space))
    _bspmm_check(C, g["out_C"], br, bc, K, g["A_vals"].dtype)


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("br,bc,K,dtype", [(32, 32, 128, "float32"), (16, 16, 128, "float32"), (32, 32, 200, "float32"),
                                           (16, 16, 36, "float32"), (32, 32, 64, "float64"), (8, 8, 50, "float32"),
                                           (32, 32, 130, "float32"), (5, 3, 17, "float64")])
def test_oracle_bspmm(space, br, bc, K, dtype):
    w = synth.make("bspmm", None, Mb=300, deg=7, br=br, bc=bc, K=K, dtype=dtype)
    w["A_pos"] = w["A_pos"].copy()
    w["A_pos"][41:] -= 7                   # block row 40 becomes empty (its blocks are dropped below)
    w["A_crd"] = np.delete(w["A_crd"], slice(40 * 7, 41 * 7))
    w["A_vals"] = np.delete(w["A_vals"].reshape(-1, br, bc), slice(40 * 7, 41 * 7), axis=0).reshape(-1)
    Mb = 300
    C = G.run("bspmm", place(w, space))
    want = oracle.bspmm(w["A_pos"], w["A_crd"], w["A_vals"].reshape(-1, br, bc), w["B"].reshape(Mb * bc, K), br, bc)
    assert not want[40 * br:41 * br].any()
    _bspmm_check(C, want, br, bc, K, dtype)


def test_bspmm_tensor_core_path_is_accurate_on_signed_data():
    # mixed-sign operands (cancellation): the 3xTF32 split must stay within 1e-5 of the fp32 result's scale
    rng = np.random.default_rng(5)
    Mb, Nb, br, bc, K = 64, 80, 32, 32, 128
    A = (rng.standard_normal((Mb * br, Nb * bc)) * (rng.random((Mb, 1, Nb, 1)) < 0.2).repeat(br, 1).repeat(bc, 3).reshape(Mb * br, Nb * bc)).astype(np.float32)
    p, c, v = formats.bcsr_from_dense(A, br, bc)
    B = rng.standard_normal((Nb * bc, K)).astype(np.float32)
    C = G.run("bspmm", dict(dims=[Mb, Nb, br, bc, K], A_pos=p, A_crd=c, A_vals=v.reshape(-1), B=B.reshape(-1)))
    want = A.astype(np.float64) @ B.astype(np.float64)
    err = np.abs(C.reshape(Mb * br, K) - want).max() / np.abs(want).max()
    assert err < 1e-5, err


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("br,K,deg", [(32, 128, 70), (16, 256, 45), (32, 8, 3), (16, 4, 5), (32, 260, 33)])
def test_bspmm_long_block_rows_and_narrow_or_ragged_tiles(space, br, K, deg):
    # block rows with more than 32 stored blocks (the TMA warp reads block columns 32 at a time), dense operands narrower
    # than one 32-column tensor box, and a last column tile that is partly out of bounds (zero-filled by TMA)
    Mb = 24
    w = synth.make("bspmm", None, Mb=Mb, deg=min(deg, Mb), br=br, bc=br, K=K, dtype="float32")
    if deg > Mb:                           # wider than tall: more block columns than block rows
        w = synth.make("bspmm", None, Mb=96, deg=deg, br=br, bc=br, K=K, dtype="float32")
        Mb = 96
    C = G.run("bspmm", place(w, space))
    want = oracle.bspmm(w["A_pos"], w["A_crd"], w["A_vals"].reshape(-1, br, br), w["B"].reshape(Mb * br, K), br, br)
    _bspmm_check(C, want, br, br, K, "float32")


def test_bspmm_empty_and_errors():
    z = np.zeros(0, np.float32)
    C = G.run("bspmm", dict(dims=[3, 4, 16, 16, 8], A_pos=np.zeros(4, np.int32), A_crd=np.zeros(0, np.int32), A_vals=z,
                            B=np.ones(4 * 16 * 8, np.float32)))
    assert C.shape == (3 * 16 * 8,) and not C.any()
    A = tb.makeBCSR("A", [3, 4, 16, 16], np.zeros(4, np.int32), np.zeros(0, np.int32), z)
    B = tb.makeDense("B", [4, 8, 8], np.ones(4 * 8 * 8, np.float32))           # block width 8 != 16
    Ct = tb.Tensor("C", [3, 16, 8], tb.Format([tb.dense] * 3), np.float32)
    k = tb.compile(G.EXPR["bspmm"], Ct, A, B)
    with pytest.raises(tb.TacoError):
        k(Ct, A, B)


# ---------------------------------------------------------------------------------------------------------
# pack(): COO -> CSR / DCSR / CSF on the device, src/tensor.cpp:295-463
# ---------------------------------------------------------------------------------------------------------
_PACK_FMT = {"csr": tb.CSR, "dcsr": tb.DCSR, "csf3": tb.CSF3, "csc": tb.Format([tb.dense, tb.compressed], [1, 0])}


def _pack_levels(t, kind):
    out = {"A_vals": G.to_host(t.vals())}
    for l, c in enumerate(t.format.levels):
        if c == tb.compressed:
            pos, crd = t.level(l)
            out[f"A{l + 1}_pos"], out[f"A{l + 1}_crd"] = G.to_host(pos), G.to_host(crd)
    return out


def _run_pack(kind, dims, coords, vals, space):
    tb.set_result_space(space)
    try:
        if space == "device":
            import torch
            coords = [torch.as_tensor(c).cuda() for c in coords]
            vals = torch.as_tensor(vals).cuda()
        t = tb.pack("A", dims, _PACK_FMT[kind], coords, vals)
        if space == "device":
            tb.synchronize()
        return {k: np.array(v) for k, v in _pack_levels(t, kind).items()}
    finally:
        tb.set_result_space("host")


@pytest.mark.parametrize("space", SPACES)
def test_kat_pack_rua32(space):
    # the reference's storage KAT for test/data/rua_32.mtx (tests-api.cpp:261-300): file order and a shuffled order
    pos, crd, vals = H.rua32_csr()
    rows = np.repeat(np.arange(32), np.diff(pos)).astype(np.int32)
    for order in (np.lexsort((rows, crd)), np.random.default_rng(5).permutation(126)):
        got = _run_pack("csr", [32, 32], [rows[order].copy(), crd[order].copy()], vals[order].copy(), space)
        assert np.array_equal(got["A2_pos"], pos) and np.array_equal(got["A2_crd"], crd) and np.array_equal(got["A_vals"], vals)


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("name", H.golden_cases("pack"))
def test_golden_pack(name, space):
    g = H.load_golden(name)
    kind = name.split("_")[1]
    dims = [int(x) for x in g["dims"]]
    got = _run_pack(kind, dims, [g[f"c{m}"] for m in range(len(dims))], g["vals"], space)
    outs = {k[4:]: v for k, v in g.items() if k.startswith("out_")}
    assert set(got) == set(outs)
    for k, want in outs.items():
        assert np.array_equal(got[k], want), k


@pytest.mark.parametrize("space", SPACES)
@pytest.mark.parametrize("kind,dims,n", [("csr", (100_000, 70_000), 1_500_000), ("dcsr", (1 << 20, 1 << 18), 300_000),
                                         ("csf3", (3_000, 500, 2_000), 2_000_000), ("csr", (50, 40), 100_000), ("csc", (300, 500), 90_000),
                                         ("csf3", (7, 5, 3), 4_000)])
def test_oracle_pack(space, kind, dims, n):
    # unsorted coordinates with duplicates (heavily duplicated in the small-dimension cases); integer values keep the
    # duplicate sums exact in any order
    rng = np.random.default_rng(n + len(dims))
    coords = [rng.integers(0, d, n).astype(np.int32) for d in dims]
    vals = np.floor(rng.random(n) * 5 + 1)
    got = _run_pack(kind, list(dims), coords, vals, space)
    want = oracle.pack(kind, list(dims), coords, vals) if n <= 100_000 else None
    if want is None:       # large cases: numpy reference without the python loop over runs
        flat = np.ravel_multi_index(coords, dims)
        uq, inv = np.unique(flat, return_inverse=True)
        sums = np.bincount(inv, weights=vals)
        uc = np.unravel_index(uq, dims)
        assert np.array_equal(got["A_vals"], sums)
        last = f"A{len(dims)}_crd"
        assert np.array_equal(got[last], uc[-1].astype(np.int32))
        if kind == "csr":
            pos = np.zeros(dims[0] + 1, np.int64)
            np.add.at(pos, uc[0] + 1, 1)
            assert np.array_equal(got["A2_pos"], np.cumsum(pos))
        else:
            assert np.array_equal(got["A1_crd"], np.unique(uc[0]))
            assert got["A1_pos"].tolist() == [0, np.unique(uc[0]).size]
            assert got[f"A{len(dims)}_pos"][-1] == uq.size
    else:
        assert set(got) == set(want)
        for k in want:
            assert np.array_equal(got[k], want[k]), k


def test_pack_then_compute_and_edge_cases():
    # a packed CSR tensor feeds SpMV directly; empty input; unsupported target
    rng = np.random.default_rng(3)
    n, m, e = 5_000, 4_000, 60_000
    r, c = rng.integers(0, n, e).astype(np.int32), rng.integers(0, m, e).astype(np.int32)
    v = np.floor(rng.random(e) * 7 + 1)
    A = tb.pack("A", [n, m], tb.CSR, [r, c], v)
    x = np.floor(rng.random(m) * 5)
    xt = tb.makeDense("x", [m], x)
    y = tb.Tensor("y", [n], tb.Format([tb.dense]), np.float64)
    tb.compile(G.EXPR["spmv"], y, A, xt)(y, A, xt)
    D = np.zeros((n, m))
    np.add.at(D, (r, c), v)
    assert np.array_equal(G.to_host(y.vals()), D @ x)
    z = np.zeros(0, np.int32)
    E = tb.pack("E", [6, 5], tb.CSR, [z, z], np.zeros(0))
    assert np.array_equal(G.to_host(E.level(1)[0]), np.zeros(7, np.int32))
    E = tb.pack("E", [6, 5, 4], tb.CSF3, [z, z, z], np.zeros(0))
    assert G.to_host(E.level(0)[0]).tolist() == [0, 0]
    with pytest.raises(tb.TacoError):
        tb.pack("B", [6, 5], tb.Format([tb.dense, tb.dense]), [z, z], np.zeros(0))


# ---------------------------------------------------------------------------------------------------------
# the _shim_ entry points (void** parameterPack, codegen_cuda.cpp:1500-1540) are CALLED, not only exported
# ---------------------------------------------------------------------------------------------------------
def test_shims_are_called_through_the_packed_convention(tmp_path):
    import ctypes
    from taco_b200 import _lib
    w = synth.make("spmv", None, n=20_011, deg=9)
    y = tb.Tensor("y", [20_011], tb.Format([tb.dense]), np.float64)
    A = tb.makeCSR("A", [20_011, 20_011], w["A_pos"], w["A_crd"], w["A_vals"])
    x = tb.makeDense("x", [20_011], w["x"])
    pack = (ctypes.c_void_p * 3)(*[ctypes.cast(t.ptr, ctypes.c_void_p) for t in (y, A, x)])
    tb.set_result_space("host")
    _lib.check(_lib.lib._shim_taco_b200_spmv_assemble(pack))
    y.adopt_results()
    _lib.check(_lib.lib._shim_taco_b200_spmv_compute(pack))
    assert np.array_equal(y.vals(), oracle.spmv(w["A_pos"], w["A_crd"], w["A_vals"], w["x"]))
    # sparse result through the shim of evaluate
    s = synth.make("spadd", None, n=5_003, deg=7)
    C = tb.Tensor("C", [5_003, 5_003], tb.CSR, np.float64)
    Aa = tb.makeCSR("A", [5_003, 5_003], s["A_pos"], s["A_crd"], s["A_vals"])
    Bb = tb.makeCSR("B", [5_003, 5_003], s["B_pos"], s["B_crd"], s["B_vals"])
    pack = (ctypes.c_void_p * 3)(*[ctypes.cast(t.ptr, ctypes.c_void_p) for t in (C, Aa, Bb)])
    _lib.check(_lib.lib._shim_taco_b200_spadd_evaluate(pack))
    C.adopt_results()
    cp, cc, cv = oracle.spadd(s["A_pos"], s["A_crd"], s["A_vals"], s["B_pos"], s["B_crd"], s["B_vals"])
    assert np.array_equal(C.level(1)[0], cp) and np.array_equal(C.level(1)[1], cc) and np.array_equal(C.vals(), cv)
    # _shim_taco_b200_read: (path, tensor)
    path = tmp_path / "m.mtx"
    path.write_text("%%MatrixMarket matrix coordinate real general\n3 4 3\n1 1 2.5\n3 4 -1e-3\n2 2 7\n")
    T = tb.Tensor("T", [0, 0], tb.CSR, np.float64)
    bpath = ctypes.create_string_buffer(str(path).encode())
    pack2 = (ctypes.c_void_p * 2)(ctypes.cast(bpath, ctypes.c_void_p), ctypes.cast(T.ptr, ctypes.c_void_p))
    _lib.check(_lib.lib._shim_taco_b200_read(pack2))
    T.dims = [int(T._dims[0]), int(T._dims[1])]
    T.adopt_results()
    assert T.dims == [3, 4] and T.level(1)[0].tolist() == [0, 1, 2, 3] and T.level(1)[1].tolist() == [0, 1, 3]
    assert T.vals().tolist() == [2.5, 7.0, -1e-3]
