"""Pins the CPU oracle (oracle/taco_oracle.c) -- `-m "not gpu"`.

1. against the reference's own known-answer vectors (test/tests-expr_storage.cpp) over its fixtures
   (test/test_tensors.cpp), restated in tests/helpers.py with their line numbers;
2. against outputs of the reference itself (tests/golden/*.npz, produced by tests/golden/make_golden.py from
   oracle/_ref built out of /root/reference).
Structure (pos/crd) must be bit-exact; values bit-exact for the integer-valued cases and for every kernel whose
operation order the oracle restates (all of them under the default schedule); north-star tolerance otherwise.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import oracle  # noqa: E402
from taco_b200 import formats  # noqa: E402
import helpers as H  # noqa: E402


# ---------------------------------------------------------------------------------------------------------
# 1. reference known-answer tests
# ---------------------------------------------------------------------------------------------------------
def test_kat_spmv():
    # tests-expr_storage.cpp:873-911: d33a("B",{Dense,Sparse})(i,k) * d3b("c")(k) == {0,0,18}
    p, c, v = formats.csr_from_dense(H.d33a())
    assert np.array_equal(oracle.spmv(p, c, v, H.d3b()), [0, 0, 18])


def test_kat_matrix_add_csr():
    # tests-expr_storage.cpp:560-577: pos {0,2,2,5}, crd {0,1,0,1,2}, vals {10,22,3,30,4}
    ap, ac, av = formats.csr_from_dense(H.d33a())
    bp, bc, bv = formats.csr_from_dense(H.d33b())
    cp, cc, cv = oracle.spadd(ap, ac, av, bp, bc, bv)
    assert cp.tolist() == [0, 2, 2, 5] and cc.tolist() == [0, 1, 0, 1, 2] and cv.tolist() == [10, 22, 3, 30, 4]


def test_kat_matrix_add_3x4():
    # tests-expr_storage.cpp:610-627: pos {0,3,3,6}, crd {0,2,3,0,2,3}, vals {4,3,3,8,5,5}
    ap, ac, av = formats.csr_from_dense(H.d34a())
    bp, bc, bv = formats.csr_from_dense(H.d34b())
    cp, cc, cv = oracle.spadd(ap, ac, av, bp, bc, bv)
    assert cp.tolist() == [0, 3, 3, 6] and cc.tolist() == [0, 2, 3, 0, 2, 3] and cv.tolist() == [4, 3, 3, 8, 5, 5]


def test_kat_matrix_mul():
    # tests-expr_storage.cpp:996-1035: d33a * d33b == {0,0,0, 0,0,0, 30,180,0}  (dense result; here via SpMM and SpGEMM)
    ap, ac, av = formats.csr_from_dense(H.d33a())
    C = oracle.spmm(ap, ac, av, H.d33b())
    assert np.array_equal(C, [[0, 0, 0], [0, 0, 0], [30, 180, 0]])
    bp, bc, bv = formats.csr_from_dense(H.d33b())
    cp, cc, cv = oracle.spgemm(ap, ac, av, bp, bc, bv, 3)
    assert np.array_equal(formats.csr_to_dense(3, 3, cp, cc, cv), C)
    assert cp.tolist() == [0, 0, 0, 2] and cc.tolist() == [0, 1]


def test_kat_mttkrp():
    # tests-expr_storage.cpp:1145-1166: d233a(i,k,l) * d33a(k,j) * d33b(l,j) == {0,80,0, 180,0,0}
    t = formats.coo_to_csf3(*H.d233a())
    A = oracle.mttkrp(t, H.d33a(), H.d33b(), 2)
    assert np.array_equal(A, [[0, 80, 0], [180, 0, 0]])


def test_kat_parafac_mttkrp2_mttkrp3():
    # tests-parafac.cpp:157-187 with the factories of test/expr_factory.cpp:100-124: the mode-J / mode-K MTTKRPs over d333a.
    # Stored in mode orderings {1,0,2} / {2,0,1}, the level arrays ARE the CSF of the permuted tensor, so each is the
    # standard MTTKRP over that storage order: A(i,r) = B(k,i,l) C(k,r) D(l,r) and A(i,r) = B(k,l,i) C(k,r) D(l,r).
    a, b, c, v = H.d333a()
    A2 = oracle.mttkrp(formats.coo_to_csf3(b, a, c, v), H.d33a(), H.d33b(), 3)                  # B'(i,k,l) = B(k,i,l)
    want2 = np.zeros((3, 3)); want2[0, 1] = 80; want2[2, 1] = 240
    assert np.array_equal(A2, want2)
    A3 = oracle.mttkrp(formats.coo_to_csf3(c, a, b, v), H.d33a(), H.d33b(), 3)                  # B''(i,k,l) = B(k,l,i)
    want3 = np.zeros((3, 3)); want3[0, 1] = 80; want3[1, 1] = 120; want3[2, 1] = 240
    assert np.array_equal(A3, want3)


def test_kat_tensor_vector_mul():
    # tests-expr_storage.cpp:1056-1072: d333a(i,j,k) * d3b(k) == {4,0,12, 0,0,33, 0,24,0}
    t = formats.coo_to_csf3(*H.d333a())
    A = oracle.ttv(t, H.d3b(), 3, 3)
    assert np.array_equal(A, [[4, 0, 12], [0, 0, 33], [0, 24, 0]])


def test_kat_tensor_matrix_mul():
    # tests-expr_storage.cpp:1114-1139: d233a(i,m,l) * d33a(l,j) == {0,4,0, 0,0,0, 12,0,16,  0,0,0, 0,0,0, 21,12,28}
    t = formats.coo_to_csf3(*H.d233a())
    A = oracle.ttm(t, H.d33a(), 2, 3)
    assert np.array_equal(A.reshape(-1), [0, 4, 0, 0, 0, 0, 12, 0, 16, 0, 0, 0, 0, 0, 0, 21, 12, 28])


def test_kat_bspmv():
    # tests-expr_storage.cpp:939-960: d3322a("B",{Dense,Sparse,Dense,Dense})(i,k,j,l) * d32b("c")(k,l)
    #                                  == {88.2, 96.4, 0.0, 0.0, 319.4, 335.8}
    pos, crd, blocks = H.d3322a()
    a = oracle.bspmv(pos, crd, blocks, H.d32b(), 2, 2)
    H.assert_close(a.reshape(-1), np.array([88.2, 96.4, 0.0, 0.0, 319.4, 335.8]), np.float64)
    # the same statement as a one-column blocked SpMM
    C = oracle.bspmm(pos, crd, blocks, H.d32b().reshape(6, 1), 2, 2)
    H.assert_close(C.reshape(-1), a.reshape(-1), np.float64)      # (different summation order: scalar temporary)


# ---------------------------------------------------------------------------------------------------------
# 2. outputs of the reference itself
# ---------------------------------------------------------------------------------------------------------
def _check_vals(name, got, want):
    if "_int_" in name or name.endswith("_default") or "_cpu" not in name:
        # integer-valued data, or the default schedule whose operation order the oracle restates: bit-exact
        assert np.array_equal(got, want), f"{name}: not bit-exact, max diff {np.max(np.abs(got - want))}"
    else:
        H.assert_close(got, want, want.dtype)


@pytest.mark.parametrize("name", H.golden_cases("spmv"))
def test_golden_spmv(name):
    g = H.load_golden(name)
    _check_vals(name, oracle.spmv(g["A_pos"], g["A_crd"], g["A_vals"], g["x"]), g["out_y"])


@pytest.mark.parametrize("name", H.golden_cases("spmm"))
def test_golden_spmm(name):
    g = H.load_golden(name)
    n, m, K = g["dims"]
    C = oracle.spmm(g["A_pos"], g["A_crd"], g["A_vals"], g["B"].reshape(m, K))
    _check_vals(name, C.reshape(-1), g["out_C"])


@pytest.mark.parametrize("name", H.golden_cases("dcsr_spmm"))
def test_golden_spmm_dcsr(name):
    g = H.load_golden(name)
    n, m, K = [int(x) for x in g["dims"]]
    C = oracle.spmm_dcsr(n, g["A1_pos"], g["A1_crd"], g["A2_pos"], g["A2_crd"], g["A_vals"], g["B"].reshape(m, K))
    _check_vals(name, C.reshape(-1), g["out_C"])
    # the same product through the CSR restatement on the expanded pos array (what the GPU path does)
    pos = formats.dcsr_to_csr(n, g)
    assert np.array_equal(oracle.spmm(pos, g["A2_crd"], g["A_vals"], g["B"].reshape(m, K)), C)


@pytest.mark.parametrize("name", H.golden_cases("densesddmm"))
def test_golden_sddmm_dense(name):
    g = H.load_golden(name)
    n, m, J = [int(x) for x in g["dims"]]
    A = oracle.sddmm_dense(g["B_pos"], g["B_crd"], g["B_vals"], g["C"].reshape(n, J), g["D"].reshape(J, m))
    _check_vals(name, A.reshape(-1), g["out_A"])


@pytest.mark.parametrize("name", H.golden_cases("sddmm"))
def test_golden_sddmm(name):
    g = H.load_golden(name)
    n, m, K = g["dims"]
    ap, ac, av = oracle.sddmm(g["B_pos"], g["B_crd"], g["B_vals"], g["C"].reshape(n, K), g["D"].reshape(m, K))
    assert np.array_equal(ap, g["out_A_pos"]) and np.array_equal(ac, g["out_A_crd"])
    _check_vals(name, av, g["out_A_vals"])


@pytest.mark.parametrize("name", H.golden_cases("mttkrp"))
def test_golden_mttkrp(name):
    g = H.load_golden(name)
    I, K, L, R = g["dims"]
    A = oracle.mttkrp(g, g["C"].reshape(K, R), g["D"].reshape(L, R), I)
    _check_vals(name, A.reshape(-1), g["out_A"])


@pytest.mark.parametrize("name", H.golden_cases("ttv"))
def test_golden_ttv(name):
    g = H.load_golden(name)
    I, K, L = g["dims"]
    _check_vals(name, oracle.ttv(g, g["c"], I, K).reshape(-1), g["out_A"])


@pytest.mark.parametrize("name", H.golden_cases("ttm"))
def test_golden_ttm(name):
    g = H.load_golden(name)
    I, K, L, R = g["dims"]
    _check_vals(name, oracle.ttm(g, g["C"].reshape(L, R), I, K).reshape(-1), g["out_A"])


@pytest.mark.parametrize("name", H.golden_cases("spadd"))
def test_golden_spadd(name):
    g = H.load_golden(name)
    cp, cc, cv = oracle.spadd(g["A_pos"], g["A_crd"], g["A_vals"], g["B_pos"], g["B_crd"], g["B_vals"])
    assert np.array_equal(cp, g["out_C_pos"]) and np.array_equal(cc, g["out_C_crd"]), "structure must be bit-exact"
    assert np.array_equal(cv, g["out_C_vals"])
    if "_frac_" in name:
        assert (cv == 0).any(), "fixture is meant to contain explicit zeros that must be kept"


@pytest.mark.parametrize("name", H.golden_cases("spgemm"))
def test_golden_spgemm(name):
    g = H.load_golden(name)
    cp, cc, cv = oracle.spgemm(g["A_pos"], g["A_crd"], g["A_vals"], g["B_pos"], g["B_crd"], g["B_vals"], int(g["dims"][2]))
    assert np.array_equal(cp, g["out_C_pos"]) and np.array_equal(cc, g["out_C_crd"]), "structure must be bit-exact"
    assert np.array_equal(cv, g["out_C_vals"])


@pytest.mark.parametrize("name", H.golden_cases("bspmv"))
def test_golden_bspmv(name):
    g = H.load_golden(name)
    Mb, Nb, br, bc = g["dims"]
    a = oracle.bspmv(g["A_pos"], g["A_crd"], g["A_vals"].reshape(-1, br, bc), g["c"].reshape(Nb, bc), br, bc)
    assert np.array_equal(a.reshape(-1), g["out_a"]), "default schedule: the oracle restates the operation order"


@pytest.mark.parametrize("name", H.golden_cases("bspmm"))
def test_golden_bspmm(name):
    g = H.load_golden(name)
    Mb, Nb, br, bc, K = g["dims"]
    C = oracle.bspmm(g["A_pos"], g["A_crd"], g["A_vals"].reshape(-1, br, bc), g["B"].reshape(Nb * bc, K), br, bc)
    assert np.array_equal(C.reshape(-1), g["out_C"]), "default schedule: the oracle restates the operation order"


@pytest.mark.parametrize("name", H.golden_cases("pack"))
def test_golden_pack(name):
    # COO -> CSR / DCSR / CSF: the oracle's restatement of TensorBase::pack() against the reference's own insert()+pack()
    g = H.load_golden(name)
    kind = name.split("_")[1]
    dims = [int(x) for x in g["dims"]]
    coords = [g[f"c{m}"] for m in range(len(dims))]
    got = oracle.pack(kind, dims, coords, g["vals"])
    outs = {k[4:]: v for k, v in g.items() if k.startswith("out_")}
    assert set(got) == set(outs), (sorted(got), sorted(outs))
    for k, want in outs.items():
        if k.endswith("_pos") or k.endswith("_crd"):
            assert np.array_equal(got[k], want), k
    # integer-valued cases (with duplicates) are exact in any order; fractional cases have distinct coordinates
    assert np.array_equal(got["A_vals"], outs["A_vals"])


def test_kat_pack_rua32():
    # tests-api.cpp:261-300: the coordinates of rua_32.mtx in file (column-major) order and in a shuffled order both pack to
    # exactly the CSR arrays the reference's own storage test expects
    pos, crd, vals = H.rua32_csr()
    assert vals[0] == 101.0 and vals[5] == 126.0 and vals[-1] == 3232.0 and pos[-1] == crd.size == 126
    rows = np.repeat(np.arange(32), np.diff(pos)).astype(np.int32)
    for order in (np.lexsort((rows, crd)), np.random.default_rng(5).permutation(126)):
        got = oracle.pack("csr", [32, 32], [rows[order], crd[order]], vals[order])
        assert np.array_equal(got["A2_pos"], pos) and np.array_equal(got["A2_crd"], crd) and np.array_equal(got["A_vals"], vals)
