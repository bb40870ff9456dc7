"""Parity at BASELINE.json's FULL sizes (`-m gpu`): the oracle cannot run 10^8..10^9 operations in seconds, so each config is
checked two ways through the C ABI on device-resident operands:

1. slab: the first R rows / slices of the full-size result against the C oracle run on exactly that slab of the operands
   (rows and slices are independent) -- bit-exact wherever the kernels keep the reference's order;
2. size-independent properties over the WHOLE result: a checksum recomputed with an independent formulation (torch gathers +
   index_add / matmul in fp64), cross-kernel identities (a column of SpMM is an SpMV), structure invariants of sparse outputs
   (pos monotone, columns strictly ascending inside a row, union / product structure), linearity.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle  # noqa: E402
import gpu_util as G  # noqa: E402
import taco_b200 as tb  # noqa: E402
import synth  # noqa: E402  (tests/synth.py: workload generators)

pytestmark = pytest.mark.gpu


def _torch():
    import torch
    return torch


def _make(name, **over):
    """full-size operands generated on the device; torch's stream is drained before the library (which runs on its own
    stream -- a stream-ordered API expects complete operands) sees them"""
    w = synth.make(name, "cuda", **over)
    _torch().cuda.synchronize()
    return w


def _rows_of(pos, nnz):
    """row id of every stored position (device)"""
    torch = _torch()
    lens = (pos[1:] - pos[:-1]).to(torch.int64)
    return torch.repeat_interleave(torch.arange(lens.numel(), device=pos.device), lens, output_size=nnz)


def _host(w, keys, rows, pos_key):
    """the first `rows` rows of a CSR operand as host arrays (pos_key = 'A' or 'B' prefix)"""
    pos = G.to_host(w[pos_key + "_pos"][: rows + 1])
    nz = int(pos[-1])
    return pos, G.to_host(w[pos_key + "_crd"][:nz]), G.to_host(w[pos_key + "_vals"][:nz])


def _rel(a, b):
    torch = _torch()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


# ---------------------------------------------------------------------------------------------------------------
def test_full_c1_spmv():
    torch = _torch()
    w = _make("spmv")                       # 1M x 1M, 10M nnz, fp64
    n, m = w["dims"]
    y = torch.as_tensor(G.run("spmv", w))
    # slab: first 200k rows bit-exact
    R = 200_000
    p, c, v = _host(w, None, R, "A")
    assert np.array_equal(y[:R].numpy(), oracle.spmv(p, c, v, G.to_host(w["x"])))
    # whole result against an independent formulation (atomic index_add, different order)
    rows = _rows_of(w["A_pos"], int(w["A_crd"].shape[0]))
    ref = torch.zeros(n, dtype=torch.float64, device="cuda").index_add_(0, rows, w["A_vals"] * w["x"][w["A_crd"].long()])
    assert _rel(y.cuda(), ref) < 1e-12
    # linearity: A (2x + z) = 2 A x + A z   (z = reversed x)
    z = torch.flip(w["x"], [0]).contiguous()
    x2 = (2 * w["x"] + z).contiguous()
    torch.cuda.synchronize()
    yz = torch.as_tensor(G.run("spmv", dict(w, x=z)))
    y2 = torch.as_tensor(G.run("spmv", dict(w, x=x2)))
    assert _rel(y2, 2 * y + yz) < 1e-12


def test_full_c2_spmm():
    torch = _torch()
    w = _make("spmm")                       # R-MAT 4M x 4M, 67M nnz, K = 128, fp32
    n, m, K = w["dims"]
    tb.set_result_space("device")
    try:
        k, ts = G.build("spmm", w)
        out = torch.empty(n * K, dtype=torch.float32, device="cuda")
        ts[0].set_vals(out)
        k.compute(*ts)
        tb.synchronize()
    finally:
        tb.set_result_space("host")
    C = out.view(n, K)
    lens = (w["A_pos"][1:] - w["A_pos"][:-1])
    # slab: first 100k rows; non-hub rows bit-exact, hub rows within the fp32 tolerance
    R = 100_000
    p, c, v = _host(w, None, R, "A")
    want = oracle.spmm(p, c, v, G.to_host(w["B"]).reshape(m, K))
    got = C[:R].cpu().numpy()
    short = np.diff(p) <= 128
    assert np.array_equal(got[short], want[short])
    # hub rows (up to 65k terms): a sequential fp32 sum is itself only accurate to ~n*eps, so both results are measured
    # against an fp64 evaluation -- ours must be within the north-star tolerance of it, or at least as close as the oracle
    exact = oracle.spmm(p, c, v.astype(np.float64), G.to_host(w["B"]).reshape(m, K).astype(np.float64))
    err_ours, err_ref = np.abs(got - exact), np.abs(want - exact)
    assert (err_ours <= np.maximum(err_ref, 1e-5 * np.abs(exact))).all()
    # cross-kernel identity: column k of SpMM is the SpMV with column k of B (same accumulation order -> identical bits)
    for col in (0, 77):
        x = w["B"].view(m, K)[:, col].contiguous()
        torch.cuda.synchronize()
        y = torch.as_tensor(G.run("spmv", dict(dims=[n, m], A_pos=w["A_pos"], A_crd=w["A_crd"], A_vals=w["A_vals"], x=x))).cuda()
        ok = (lens <= 128)
        assert torch.equal(C[:, col][ok], y[ok])
        assert _rel(C[:, col].double(), y.double()) < 1e-5
    # checksum of the whole result: 1^T C = (1^T A) B, recomputed in fp64
    colw = torch.zeros(m, dtype=torch.float64, device="cuda").index_add_(0, w["A_crd"].long(), w["A_vals"].double())
    ref = colw @ w["B"].view(m, K).double()
    assert _rel(C.double().sum(0), ref) < 1e-5


def test_full_c3_sddmm():
    torch = _torch()
    w = _make("sddmm")                      # 2M x 2M, 40M nnz, K = 64, fp32
    n, m, K = w["dims"]
    pos, crd, vals = G.run("sddmm", w)
    assert np.array_equal(pos, G.to_host(w["B_pos"])) and np.array_equal(crd, G.to_host(w["B_crd"]))     # structure of B
    # slab: first 50k rows against the oracle (lane-parallel dot product: fp32 tolerance)
    R = 50_000
    p, c, v = _host(w, None, R, "B")
    _, _, want = oracle.sddmm(p, c, v, G.to_host(w["C"][: R * K]).reshape(R, K), G.to_host(w["D"]).reshape(m, K))
    assert np.allclose(vals[: p[-1]], want, rtol=1e-5, atol=0)
    # a random sample of 2M nonzeros of the whole result, recomputed in fp64 with torch gathers
    g = torch.Generator(device="cuda").manual_seed(3)
    s = torch.randint(0, int(crd.shape[0]), (2_000_000,), device="cuda", generator=g)
    rows = _rows_of(w["B_pos"], int(crd.shape[0]))[s]
    ref = w["B_vals"][s].double() * (w["C"].view(n, K)[rows].double() * w["D"].view(m, K)[w["B_crd"][s].long()].double()).sum(1)
    got = torch.as_tensor(vals).cuda()[s].double()
    assert float(((got - ref).abs() / ref.abs()).max()) < 1e-5


def test_full_c4_mttkrp():
    torch = _torch()
    w = _make("mttkrp")                     # 10M x 1M x 1M, 200M nnz, R = 32, fp64
    I, K, L, R = w["dims"]
    tb.set_result_space("device")
    try:
        k, ts = G.build("mttkrp", w)
        out = torch.empty(I * R, dtype=torch.float64, device="cuda")
        ts[0].set_vals(out)
        k.compute(*ts)
        tb.synchronize()
    finally:
        tb.set_result_space("host")
    A = out.view(I, R)
    # slab: the first 100k slices bit-exact against the oracle (rows without a slice are zero)
    S = 100_000
    p2 = G.to_host(w["B2_pos"][: S + 1]); nf = int(p2[-1])
    p3 = G.to_host(w["B3_pos"][: nf + 1]); nz = int(p3[-1])
    t = dict(B1_pos=np.array([0, S], np.int32), B1_crd=G.to_host(w["B1_crd"][:S]), B2_pos=p2, B2_crd=G.to_host(w["B2_crd"][:nf]),
             B3_pos=p3, B3_crd=G.to_host(w["B3_crd"][:nz]), B_vals=G.to_host(w["B_vals"][:nz]))
    last = int(t["B1_crd"][-1]) + 1
    want = oracle.mttkrp(t, G.to_host(w["C"]).reshape(K, R), G.to_host(w["D"]).reshape(L, R), last)
    assert np.array_equal(A[:last].cpu().numpy(), want)
    # checksum of the whole result: sum_i A(i,j) = sum over leaves v * C(k,j) * D(l,j), recomputed in chunks with torch
    nnz = int(w["B3_crd"].shape[0])
    nfib = int(w["B2_crd"].shape[0])
    ref = torch.zeros(R, dtype=torch.float64, device="cuda")
    Cm, Dm = w["C"].view(K, R), w["D"].view(L, R)
    step = 2_000_000                                      # fibers per chunk
    for f0 in range(0, nfib, step):
        f1 = min(f0 + step, nfib)
        p = w["B3_pos"][f0: f1 + 1].long()
        kk = torch.repeat_interleave(w["B2_crd"][f0:f1].long(), p[1:] - p[:-1])
        a, b = int(p[0]), int(p[-1])
        ref += ((w["B_vals"][a:b, None] * Cm[kk]) * Dm[w["B3_crd"][a:b].long()]).sum(0)
    assert _rel(A.sum(0), ref) < 1e-10
    assert nnz == int(w["B3_pos"][-1])


@pytest.mark.parametrize("fam", ["spadd", "spgemm"])
def test_full_c5_sparse_output(fam):
    torch = _torch()
    w = _make(fam)                          # 1M x 1M, 10M nnz each, fp64
    n = w["dims"][0]
    tb.set_result_space("device")
    try:
        k, ts = G.build(fam, w)
        k(*ts)
        tb.synchronize()
        res = ts[0]
        pos, crd = res.level(1)
        vals = res.vals()
        pos, crd, vals = pos.clone(), crd.clone(), vals.clone()
    finally:
        tb.set_result_space("host")
    nnzC = int(pos[-1])
    assert int(pos[0]) == 0 and bool((pos[1:] >= pos[:-1]).all()) and crd.numel() == nnzC == vals.numel()
    rows = _rows_of(pos, nnzC)
    same_row = rows[1:] == rows[:-1]
    assert bool((crd[1:][same_row] > crd[:-1][same_row]).all()), "columns must be strictly ascending inside a row"
    # slab: first 100k rows bit-exact (structure and values) against the oracle on the slab
    R = 100_000
    ap, ac, av = _host(w, None, R, "A")
    if fam == "spadd":
        bp, bc, bv = _host(w, None, R, "B")
        cp, cc, cv = oracle.spadd(ap, ac, av, bp, bc, bv)
    else:
        cp, cc, cv = oracle.spgemm(ap, ac, av, G.to_host(w["B_pos"]), G.to_host(w["B_crd"]), G.to_host(w["B_vals"]), n)
    e = int(cp[-1])
    assert np.array_equal(pos[: R + 1].cpu().numpy(), cp) and np.array_equal(crd[:e].cpu().numpy(), cc)
    assert np.array_equal(vals[:e].cpu().numpy(), cv)
    # checksum of the whole result
    if fam == "spadd":
        assert _rel(vals.sum()[None], (w["A_vals"].sum() + w["B_vals"].sum())[None]) < 1e-12
        # |C| = |A| + |B| - |A and B|: every key of A and of B appears exactly once
        keyC = rows * n + crd.long()
        keyA = _rows_of(w["A_pos"], int(w["A_crd"].shape[0])) * n + w["A_crd"].long()
        keyB = _rows_of(w["B_pos"], int(w["B_crd"].shape[0])) * n + w["B_crd"].long()
        assert torch.equal(keyC, torch.unique(torch.cat([keyA, keyB])))
    else:
        # 1^T C 1 = sum_p A(p) * rowsum_B(col(p))
        rb = torch.zeros(n, dtype=torch.float64, device="cuda").index_add_(
            0, _rows_of(w["B_pos"], int(w["B_crd"].shape[0])), w["B_vals"])
        ref = (w["A_vals"] * rb[w["A_crd"].long()]).sum()
        assert _rel(vals.sum()[None], ref[None]) < 1e-12


def test_full_bspmm_and_property():
    torch = _torch()
    w = _make("bspmm")                      # 32768^2 blocks of 32 x 32, 16 per block row, K = 128, fp32
    Mb, Nb, br, bc, K = w["dims"]
    tb.set_result_space("device")
    try:
        k, ts = G.build("bspmm", w)
        out = torch.empty(Mb * br * K, dtype=torch.float32, device="cuda")
        ts[0].set_vals(out)
        k.compute(*ts)
        tb.synchronize()
    finally:
        tb.set_result_space("host")
    C = out.view(Mb * br, K)
    # slab: first 512 block rows against the oracle (tensor-core path: fp32 tolerance on the scale of the result)
    R = 512
    pos = G.to_host(w["A_pos"][: R + 1]); nb = int(pos[-1])
    want = oracle.bspmm(pos, G.to_host(w["A_crd"][:nb]), G.to_host(w["A_vals"][: nb * br * bc]).reshape(-1, br, bc),
                        G.to_host(w["B"]).reshape(Nb * bc, K), br, bc)
    got = C[: R * br].cpu().numpy()
    assert np.abs(got - want.reshape(R * br, K)).max() <= 1e-5 * np.abs(want).max()
    # checksum of the whole result: 1^T C = (column sums of A as a dense row vector) B, in fp64
    blocks = w["A_vals"].view(-1, br, bc).double().sum(1)                       # (nnzb, bc): column sums inside each block
    colw = torch.zeros(Nb, bc, dtype=torch.float64, device="cuda").index_add_(0, w["A_crd"].long(), blocks)
    ref = colw.view(1, Nb * bc) @ w["B"].view(Nb * bc, K).double()
    assert _rel(C.double().sum(0), ref[0]) < 1e-5
