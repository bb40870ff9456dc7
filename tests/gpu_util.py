"""Drivers used by the GPU parity tests, smoke() and bench.py: build taco tensors over a workload dict (numpy =>
host buffers through the C ABI with staging; torch CUDA tensors => device-resident, zero-copy) and run a family."""
import numpy as np

import taco_b200 as tb

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

EXPR = {
    "spmv": "y(i) = A(i,j) * x(j)",
    "spmm": "C(i,k) = A(i,j) * B(j,k)",
    "spmm_dcsr": "C(i,k) = A(i,j) * B(j,k)",
    "sddmm": "A(i,j) = B(i,j) * C(i,k) * D(j,k)",
    "sddmm_dense": "A(i,k) = B(i,k) * C(i,j) * D(j,k)",
    "mttkrp": "A(i,j) = B(i,k,l) * C(k,j) * D(l,j)",
    "ttv": "A(i,j) = B(i,j,k) * c(k)",
    "ttm": "A(i,j,l) = B(i,j,k) * C(k,l)",
    "spadd": "C(i,j) = A(i,j) + B(i,j)",
    "spgemm": "C(i,k) = A(i,j) * B(j,k)",
    "bspmv": "a(i,j) = A(i,k,j,l) * c(k,l)",
    "bspmm": "C(i,j,m) = A(i,k,j,l) * B(k,l,m)",
}


def is_dev(a):
    return torch is not None and isinstance(a, torch.Tensor) and a.is_cuda


def np_dtype(a):
    if torch is not None and isinstance(a, torch.Tensor):
        return np.dtype(np.float32) if a.dtype == torch.float32 else np.dtype(np.float64)
    return a.dtype


def to_device(w):
    return {k: (torch.as_tensor(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in w.items()}


def to_host(a):
    if torch is not None and isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def empty_like_space(w_array, count, dtype):
    """result buffer in the same space as the operands (device operands => device result, no copies)"""
    if is_dev(w_array):
        return torch.empty(count, dtype=torch.float32 if np.dtype(dtype) == np.float32 else torch.float64, device="cuda")
    return np.empty(count, dtype=dtype)


class _PackHolder:
    """stands where a result Tensor stands in the (kernel, tensors) pair: holds the tensor the last pack() produced"""

    def __init__(self, dims):
        self.dims, self.t = list(dims), None

    @property
    def ct(self):
        return self.t.ct


class _PackKernel:
    """pack(): unsorted COO (w["c0"], w["c1"], w["vals"]) -> a CSR tensor (taco_b200_pack through the C ABI)"""

    def __call__(self, holder, w):
        holder.t = tb.pack("A", holder.dims, tb.CSR, [w["c0"], w["c1"]], w["vals"])
        return True


def build(family, w, colmajor_c=False):
    """returns (kernel, [result, operands...]) for a workload dict with the taco_b200.synth / tbin key names"""
    d = [int(x) for x in w["dims"]]
    if family == "spmv":
        dt = np_dtype(w["A_vals"])
        A = tb.makeCSR("A", d[:2], w["A_pos"], w["A_crd"], w["A_vals"])
        x = tb.makeDense("x", [d[1]], w["x"])
        y = tb.Tensor("y", [d[0]], tb.Format([tb.dense]), dt)
        ts = [y, A, x]
    elif family == "spmm":
        dt = np_dtype(w["A_vals"])
        A = tb.makeCSR("A", d[:2], w["A_pos"], w["A_crd"], w["A_vals"])
        B = tb.makeDense("B", [d[1], d[2]], w["B"])
        C = tb.Tensor("C", [d[0], d[2]], tb.Format([tb.dense, tb.dense], [1, 0] if colmajor_c else None), dt)
        ts = [C, A, B]
    elif family == "spmm_dcsr":
        dt = np_dtype(w["A_vals"])
        A = tb.makeDCSR("A", d[:2], w)
        B = tb.makeDense("B", [d[1], d[2]], w["B"])
        C = tb.Tensor("C", [d[0], d[2]], tb.Format([tb.dense, tb.dense], [1, 0] if colmajor_c else None), dt)
        ts = [C, A, B]
    elif family == "sddmm":
        dt = np_dtype(w["B_vals"])
        B = tb.makeCSR("B", d[:2], w["B_pos"], w["B_crd"], w["B_vals"])
        C = tb.makeDense("C", [d[0], d[2]], w["C"])
        D = tb.makeDense("D", [d[1], d[2]], w["D"])
        A = tb.Tensor("A", d[:2], tb.CSR, dt)
        ts = [A, B, C, D]
    elif family == "sddmm_dense":          # dims = (I, K, J): A, B are I x K, C is I x J, D is J x K
        dt = np_dtype(w["B_vals"])
        B = tb.makeCSR("B", d[:2], w["B_pos"], w["B_crd"], w["B_vals"])
        C = tb.makeDense("C", [d[0], d[2]], w["C"])
        D = tb.makeDense("D", [d[2], d[1]], w["D"])
        A = tb.Tensor("A", d[:2], tb.Format([tb.dense, tb.dense]), dt)
        ts = [A, B, C, D]
    elif family in ("mttkrp", "ttv", "ttm"):
        dt = np_dtype(w["B_vals"])
        B = tb.makeCSF3("B", d[:3], w)
        if family == "mttkrp":
            C = tb.makeDense("C", [d[1], d[3]], w["C"])
            D = tb.makeDense("D", [d[2], d[3]], w["D"])
            A = tb.Tensor("A", [d[0], d[3]], tb.Format([tb.dense, tb.dense]), dt)
            ts = [A, B, C, D]
        elif family == "ttv":
            c = tb.makeDense("c", [d[2]], w["c"])
            A = tb.Tensor("A", [d[0], d[1]], tb.Format([tb.dense, tb.dense]), dt)
            ts = [A, B, c]
        else:
            C = tb.makeDense("C", [d[2], d[3]], w["C"])
            A = tb.Tensor("A", [d[0], d[1], d[3]], tb.Format([tb.dense] * 3), dt)
            ts = [A, B, C]
    elif family in ("spadd", "spgemm"):
        dt = np_dtype(w["A_vals"])
        A = tb.makeCSR("A", d[:2], w["A_pos"], w["A_crd"], w["A_vals"])
        bd = d[:2] if family == "spadd" else [d[1], d[2]]
        B = tb.makeCSR("B", bd, w["B_pos"], w["B_crd"], w["B_vals"])
        C = tb.Tensor("C", [d[0], bd[1]], tb.CSR, dt)
        ts = [C, A, B]
    elif family in ("bspmv", "bspmm"):
        dt = np_dtype(w["A_vals"])
        A = tb.makeBCSR("A", d[:4], w["A_pos"], w["A_crd"], w["A_vals"])
        if family == "bspmv":
            c = tb.makeDense("c", [d[1], d[3]], w["c"])
            a = tb.Tensor("a", [d[0], d[2]], tb.Format([tb.dense, tb.dense]), dt)
            ts = [a, A, c]
        else:
            B = tb.makeDense("B", [d[1], d[3], d[4]], w["B"])
            C = tb.Tensor("C", [d[0], d[2], d[4]], tb.Format([tb.dense] * 3), dt)
            ts = [C, A, B]
    elif family == "pack":
        return _PackKernel(), [_PackHolder(d[:2]), w]
    else:
        raise KeyError(family)
    return tb.compile(EXPR[family], *ts), ts


def run(family, w, colmajor_c=False, phases="evaluate"):
    """Run a family through the C ABI.  Dense results: returns the values (numpy).  Sparse results: (pos, crd, vals).
    Device-resident operands put results in device space (then copied out here for checking)."""
    first = next(v for k, v in w.items() if k != "dims")
    dev = is_dev(first)
    tb.set_result_space("device" if dev else "host")
    try:
        k, ts = build(family, w, colmajor_c)
        if phases == "evaluate":
            k(*ts)
        else:
            k.assemble(*ts)
            k.compute(*ts)
        if dev:
            tb.synchronize()
        res = ts[0]
        if res.format.levels[-1] == tb.compressed:
            pos, crd = res.level(1)
            return to_host(pos).copy(), to_host(crd).copy(), to_host(res.vals()).copy()
        return to_host(res.vals()).copy()
    finally:
        tb.set_result_space("host")
