"""The real drop-in: the reference library with patches/module_b200.patch applied (oracle/_ref_b200/, built by
`make -C oracle ref_b200` from /root/reference) routes `Tensor::compile()` / `Kernel` / the `taco` command-line tool to
libtaco_b200.so when TACO_B200=1 -- no compileSource(), no stub file, no change to the calling program.

  * CPU (`-m "not gpu"`): the scheduled statements of the reference's GPU tests are CLASSIFIED and BOUND through the real
    lowering pipeline (lower() -> Module::compile() -> taco_b200_module_open_args): every case must get as far as the
    library call, which then refuses for want of a GPU; a GPU-scheduled statement off the hot path must be refused
    outright.  Also the commutative classifier through the C ABI.
  * GPU (`-m gpu`): oracle/b200_schedules.cpp (the reference's GPU schedules through plain compile(), results equal to the
    reference's C codegen), the reference's own test/tests-scheduling-eval.cpp `scheduling_eval.*GPU` cases compiled in
    place against the patched library, and `taco "<expr>" -cuda -s="...parallelize(...GPUBlock...)"`.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
B200 = os.path.join(ROOT, "oracle", "_ref_b200")
LIB = os.path.join(ROOT, "taco_b200", "lib", "libtaco_b200.so")
CASES = ["spmvGPU", "spmvDefaultSchedule", "spmvCommuted", "spmmGPU", "spmmDCSRGPU", "sddmmGPU", "ttvGPU", "ttmGPU",
         "mttkrpGPU", "spaddCSR", "spgemmCSR"]


def _env(tmp_path):
    return dict(os.environ, TACO_B200="1", TACO_B200_LIB=LIB, TACO_CFLAGS="-O3 -std=gnu99", TMPDIR=str(tmp_path))


def _need(name):
    exe = os.path.join(B200, name)
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref_b200/{name} not built (make -C oracle ref_b200 needs /root/reference)")
    return exe


def test_classifier_is_commutative_and_maps_the_argument_pack():
    from taco_b200 import _lib
    L = _lib.lib

    def fam(expr, fm, dt="f64", args=None):
        m = L.taco_b200_module_open_args(expr.encode(), fm.encode(), dt.encode(), args.encode() if args else None)
        return L.taco_b200_module_family(m).decode() if m else None

    assert fam("y(i) = x(j) * A(i,j)", "A:ds") == "spmv"
    assert fam("A(i,j) = B(i,j) * D(j,k) * C(i,k)", "A:ds,B:ds", "f32") == "sddmm"
    assert fam("C(i,k) = B(j,k) * A(i,j)", "A:ds") == "spmm"
    assert fam("A(i,j) = D(l,j) * B(i,k,l) * C(k,j)", "B:sss") == "mttkrp"
    assert fam("C(i,j) = B(i,j) + A(i,j)", "A:ds,B:ds,C:ds") == "spadd"
    assert fam("y(i) += A(i,j) * x(j)", "A:ds", "f64", "y,A,x") == "spmv"
    assert fam("y(i) = A(i,j) * x(j) + z(i)", "A:ds") is None            # products or sums, not mixtures
    assert fam("y(i) = A(i,j) * x(j)", "A:ds", "f64", "y,A") is None        # the argument list must name every tensor
    # a permuted module has no raw entry point (it cannot reorder arguments); the stub source and call_packed do
    m = L.taco_b200_module_open_args(b"y(i) = x(j) * A(i,j)", b"A:ds", b"f64", None)
    assert not L.taco_b200_module_get_func_ptr(m, b"compute")
    L.taco_b200_module_stub_source.restype = ctypes.c_char_p
    L.taco_b200_module_stub_source.argtypes = [ctypes.c_void_p]
    assert "fn(t0, t2, t1)" in L.taco_b200_module_stub_source(m).decode()


def test_patched_reference_binds_every_gpu_statement(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test below")
    exe = _need("b200_schedules")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=_env(tmp_path))
    out = r.stdout
    for case in CASES:
        line = next((l for l in out.splitlines() if l.startswith(case + " ")), None)
        assert line is not None, out[-3000:]
        # bound to the library: the call went through Module::callFuncPacked -> taco_b200_module_call_packed and was
        # refused there because this container has no GPU (the library has no CPU fallback)
        idx = out.index(line)
        assert "no CUDA device available" in out[idx: idx + 600], out[idx: idx + 600]
    assert "offPathStatementRefused OK" in out


@pytest.mark.gpu
def test_reference_gpu_schedules_through_plain_compile(tmp_path):
    exe = _need("b200_schedules")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=_env(tmp_path))
    for case in CASES + ["offPathStatementRefused"]:
        assert f"{case} OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    assert r.returncode == 0 and "ALL OK" in r.stdout


@pytest.mark.gpu
def test_reference_scheduling_eval_gpu_tests_against_the_library(tmp_path):
    """the reference's own gtest bodies (test/tests-scheduling-eval.cpp:1210-1587), unmodified"""
    exe = _need("taco_sched_tests")
    r = subprocess.run([exe, "--gtest_filter=scheduling_eval.*GPU*"], capture_output=True, text=True, timeout=1800, env=_env(tmp_path))
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    for t in ("spmvGPU", "spmmGPU", "spmmDCSRGPU", "sddmmGPU", "ttmGPU", "ttvGPU", "mttkrpGPU"):
        assert f"[       OK ] scheduling_eval.{t}" in r.stdout, r.stdout[-4000:]


def _write_tns(path, shape, dense, rng):
    vals = {}
    with open(path, "w") as fh:
        for idx in np.ndindex(*shape):
            last = all(i == n - 1 for i, n in zip(idx, shape))      # pins the dimensions the CLI infers from the file
            if dense or last or rng.random() < 0.15:
                v = int(rng.integers(1, 9))
                vals[idx] = v
                fh.write(" ".join(str(i + 1) for i in idx) + f" {v}\n")
    a = np.zeros(shape)
    for idx, v in vals.items():
        a[idx] = v
    return a


@pytest.mark.gpu
@pytest.mark.parametrize("expr,fmts,shapes,sched", [
    ("y(i) = A(i,j) * x(j)", dict(y="d", A="ds", x="d"), dict(A=(300, 200), x=(200,)),
     "split(i,i0,i1,32),parallelize(i0,GPUBlock,NoRaces),parallelize(i1,GPUThread,NoRaces)"),
    ("y(i) = x(j) * A(i,j)", dict(y="d", A="ds", x="d"), dict(A=(300, 200), x=(200,)), None),
    ("C(i,k) = A(i,j) * B(j,k)", dict(C="dd", A="ds", B="dd"), dict(A=(120, 90), B=(90, 16)),
     "split(i,i0,i1,32),parallelize(i0,GPUBlock,NoRaces),parallelize(i1,GPUThread,NoRaces)"),
])
def test_cli_cuda_flag_runs_on_the_library(tmp_path, expr, fmts, shapes, sched):
    """`taco "<expr>" -cuda [-s=...GPUBlock...]` on the patched CLI: kernels run in libtaco_b200, the result file equals numpy"""
    exe = _need("taco")
    rng = np.random.default_rng(7)
    args = [exe, expr, "-cuda"]
    ops = {}
    for name, f in fmts.items():
        args.append(f"-f={name}:{f}")
    for name, shape in shapes.items():
        path = tmp_path / f"{name}.tns"
        ops[name] = _write_tns(path, shape, "s" not in fmts[name], rng)
        args.append(f"-i={name}:{path}")
    if sched:
        args.append(f"-s={sched}")
    res = expr.split("(")[0].strip()
    out = tmp_path / "out.tns"
    args.append(f"-o={res}:{out}")
    r = subprocess.run(args, capture_output=True, text=True, timeout=600, env=_env(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = ops["A"] @ (ops["x"] if "x" in ops else ops["B"])
    got = np.zeros(want.shape)
    for line in open(out):
        p = line.split()
        if len(p) >= 2:
            got[tuple(int(t) - 1 for t in p[:-1])] = float(p[-1])
    assert np.array_equal(got, want)
