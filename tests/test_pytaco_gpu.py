"""pytaco-style front end with zero-copy device tensors (`-m gpu`): taco_b200.pytaco mirrors the reference's Python API
(python_bindings/pytaco/pytensor/taco_tensor.py) for the hot path; device arrays are attached, never copied."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "oracle"), HERE]
import oracle  # noqa: E402
import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def test_device_operands_are_attached_not_copied_and_results_export_dlpack():
    import torch
    from taco_b200 import pytaco as pt
    w = synth.make("spmm", None, scale=11, K=32, dtype="float32")
    n, m, K = w["dims"]
    pos, crd, vals = (torch.as_tensor(w[k]).cuda() for k in ("A_pos", "A_crd", "A_vals"))
    B = torch.as_tensor(w["B"]).cuda().view(m, K)
    A = pt.from_sp_csr((pos, crd, vals), shape=(n, m))
    Bt = pt.from_array(B)
    # zero-copy: the taco_tensor_t points at the producer's memory
    assert A._t.ct.vals == vals.data_ptr() and Bt._t.ct.vals == B.data_ptr()
    assert A.level_arrays(1)[1].data_ptr() == crd.data_ptr()
    C = pt.evaluate("C(i,k) = A(i,j) * B(j,k)", A, Bt)
    assert C.on_device() and C.shape == [n, K]
    want = oracle.spmm(w["A_pos"], w["A_crd"], w["A_vals"], w["B"].reshape(m, K))
    got = C.to_torch()
    assert got.is_cuda and got.data_ptr() == C._t.ct.vals                    # a view of the library's result, no copy
    assert np.allclose(got.cpu().numpy(), want, rtol=1e-5, atol=0)
    via_dlpack = torch.from_dlpack(C)
    assert via_dlpack.data_ptr() == got.data_ptr() and torch.equal(via_dlpack, got)
    assert C.__cuda_array_interface__["data"][0] == got.data_ptr()
    # a __cuda_array_interface__ producer that is not a torch tensor
    class Foreign:
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__
    x = torch.rand(m, dtype=torch.float32, device="cuda")
    xt = pt.from_array(Foreign(x))
    assert xt._t.ct.vals == x.data_ptr()
    y = pt.matmul(A, xt)
    assert np.allclose(y.to_array(), oracle.spmv(w["A_pos"], w["A_crd"], w["A_vals"], x.cpu().numpy()), rtol=1e-5, atol=0)


def test_host_sources_scipy_and_numpy():
    import scipy.sparse as sp
    from taco_b200 import pytaco as pt
    rng = np.random.default_rng(4)
    A = sp.random(300, 200, density=0.05, format="csr", random_state=1, dtype=np.float64)
    A.data = np.floor(A.data * 8) + 1
    A.sort_indices()
    B = np.floor(rng.random((200, 16)) * 5)
    C = pt.matmul(pt.from_sp_csr(A), pt.from_array(B))
    assert not C.on_device() and np.array_equal(C.to_array(), A @ B)
    # sparse x sparse -> sparse result (GPU assembly), exported back to scipy
    Bs = sp.random(200, 150, density=0.05, format="csr", random_state=2, dtype=np.float64)
    Bs.data = np.floor(Bs.data * 8) + 1
    Bs.sort_indices()
    G = pt.evaluate("C(i,k) = A(i,j) * B(j,k)", pt.from_sp_csr(A), pt.from_sp_csr(Bs), out_format=pt.csr)
    got = G.to_sp_csr()
    want = (A @ Bs).tocsr()
    want.sort_indices()
    assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data)
    # column-major operand through the csc constructor: y = A x with A given as CSC is not a hot-path statement -> refused
    with pytest.raises(pt.TacoError):
        pt.matmul(pt.from_sp_csc(A.tocsc()), pt.from_array(B))
