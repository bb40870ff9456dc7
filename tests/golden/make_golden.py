#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REAL reference (oracle/_ref/taco_ref_harness -> libtaco.so built from
/root/reference by oracle/Makefile) on small seeded inputs.  Run in the build container only (the GPU box has no
/root/reference and only consumes the committed .npz files):

    make -C oracle ref && python tests/golden/make_golden.py

Inputs mimic the reference's own scheduling tests (test/tests-scheduling-eval.cpp): srand-style sparse fills with
shapes taken from spmvGPU :1210, spmmGPU :1258, sddmmGPU :1360, mttkrpGPU :1526, spmataddCPU :664, spgemm :600 --
once with small-integer values (sums are exact in any order, the trick those tests rely on) and once with
fractional values (order-sensitive; compared with the north-star tolerances).
The reference's C is JIT-compiled with TACO_CFLAGS="-O3 -std=c99" (the default adds -ffast-math,
/root/reference/src/codegen/module.cpp:134).
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tbin  # noqa: E402  (oracle/tbin.py)
from taco_b200 import formats  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "taco_ref_harness")
OUT = os.path.dirname(os.path.abspath(__file__))


def run_ref(kernel, arrays, dtype, schedule):
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.tbin"), os.path.join(td, "out.tbin")
        tbin.write(fin, arrays)
        env = dict(os.environ, TACO_CFLAGS="-O3 -std=c99 -fPIC -shared", OMP_NUM_THREADS="4")
        r = subprocess.run([HARNESS, kernel, fin, fout, "--dtype", dtype, "--schedule", schedule, "--threads", "4"],
                           capture_output=True, text=True, env=env)
        if r.returncode != 0:
            raise RuntimeError(f"{kernel}/{schedule}: {r.stderr[-2000:]}")
        json.loads(r.stdout.strip().splitlines()[-1])
        return tbin.read(fout)


def sparse_fill(rng, shape, sparsity, integer, dtype):
    mask = rng.random(shape) < sparsity
    if integer:
        vals = np.floor(rng.random(shape) * 3 / max(sparsity, 0.05)).astype(dtype) + 1
    else:
        vals = (rng.random(shape) + 0.25).astype(dtype)
    return np.where(mask, vals, 0).astype(dtype)


def dense_fill(rng, shape, integer, dtype):
    if integer:
        return np.floor(rng.random(shape) * 9).astype(dtype)
    return (rng.random(shape) - 0.3).astype(dtype)


def csf_from_dense(d):
    i, k, l = np.nonzero(d)
    return formats.coo_to_csf3(i, k, l, d[i, k, l])


def main():
    cases = {}
    for integer in (True, False):
        tag = "int" if integer else "frac"
        for dtype, sfx in ((np.float64, "f64"), (np.float32, "f32")):
            rng = np.random.default_rng(94353 + (0 if integer else 1) + (0 if sfx == "f64" else 7))
            # ---- spmv (rows 5.. are empty, one long row) -------------------------------------------------------
            A = sparse_fill(rng, (102, 103), 0.08, integer, dtype)
            A[5:9] = 0
            A[40] = dense_fill(rng, 103, integer, dtype) + 1
            p, c, v = formats.csr_from_dense(A)
            x = dense_fill(rng, 103, integer, dtype)
            inp = dict(dims=np.array([102, 103], np.int32), A_pos=p, A_crd=c, A_vals=v, x=x)
            for sched in ("default", "cpu"):
                out = run_ref("spmv", inp, sfx, sched)
                cases[f"spmv_{tag}_{sfx}_{sched}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            # ---- spmm -----------------------------------------------------------------------------------------
            for K in (12, 7):
                A = sparse_fill(rng, (51, 47), 0.3, integer, dtype)
                A[7] = 0
                p, c, v = formats.csr_from_dense(A)
                B = dense_fill(rng, (47, K), integer, dtype)
                inp = dict(dims=np.array([51, 47, K], np.int32), A_pos=p, A_crd=c, A_vals=v, B=B)
                out = run_ref("spmm", inp, sfx, "default")
                cases[f"spmm_{tag}_{sfx}_K{K}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            # ---- sddmm ----------------------------------------------------------------------------------------
            Bm = sparse_fill(rng, (40, 38), 0.2, integer, dtype)
            p, c, v = formats.csr_from_dense(Bm)
            C = dense_fill(rng, (40, 16), integer, dtype)
            D = dense_fill(rng, (38, 16), integer, dtype)
            inp = dict(dims=np.array([40, 38, 16], np.int32), B_pos=p, B_crd=c, B_vals=v, C=C, D=D)
            out = run_ref("sddmm", inp, sfx, "default")
            cases[f"sddmm_{tag}_{sfx}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            # ---- mttkrp / ttv / ttm (reference GPU test shape 25 x 25 x 30, rank 32) ---------------------------
            Bt = sparse_fill(rng, (25, 25, 30), 0.1, integer, dtype)
            Bt[3] = 0
            t = csf_from_dense(Bt)
            Cm = dense_fill(rng, (25, 32), integer, dtype)
            Dm = dense_fill(rng, (30, 32), integer, dtype)
            inp = dict(dims=np.array([25, 25, 30, 32], np.int32), C=Cm, D=Dm, **t)
            out = run_ref("mttkrp", inp, sfx, "default")
            cases[f"mttkrp_{tag}_{sfx}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            cvec = dense_fill(rng, 30, integer, dtype)
            inp = dict(dims=np.array([25, 25, 30], np.int32), c=cvec, **t)
            out = run_ref("ttv", inp, sfx, "default")
            cases[f"ttv_{tag}_{sfx}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            inp = dict(dims=np.array([25, 25, 30, 8], np.int32), C=Dm[:, :8].copy(), **t)
            out = run_ref("ttm", inp, sfx, "default")
            cases[f"ttm_{tag}_{sfx}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            # ---- spadd (reference shape 1000 x 10, sparsity .15 scaled down) + spgemm (100^3, .03) -------------
            A = sparse_fill(rng, (300, 10), 0.15, integer, dtype)
            Bm = sparse_fill(rng, (300, 10), 0.15, integer, dtype)
            if not integer:
                Bm[A != 0] = np.where(rng.random(np.count_nonzero(A)) < 0.3, -A[A != 0], Bm[A != 0])  # explicit zeros in C
            ap, ac, av = formats.csr_from_dense(A)
            bp, bc, bv = formats.csr_from_dense(Bm)
            inp = dict(dims=np.array([300, 10], np.int32), A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv)
            for sched in ("default", "cpu"):
                out = run_ref("spadd", inp, sfx, sched)
                cases[f"spadd_{tag}_{sfx}_{sched}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
            A = sparse_fill(rng, (100, 100), 0.03, integer, dtype)
            Bm = sparse_fill(rng, (100, 100), 0.03, integer, dtype)
            A[10] = sparse_fill(rng, 100, 0.6, integer, dtype)       # a row with many products
            ap, ac, av = formats.csr_from_dense(A)
            bp, bc, bv = formats.csr_from_dense(Bm)
            inp = dict(dims=np.array([100, 100, 100], np.int32), A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc,
                       B_vals=bv)
            for sched in ("default", "cpu"):
                out = run_ref("spgemm", inp, sfx, sched)
                cases[f"spgemm_{tag}_{sfx}_{sched}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
    # ---- blocked SpMV / SpMM (BCSR = {Dense,Compressed,Dense,Dense}, the reference's bspmv statement) -----------
    for integer in (True, False):
        tag = "int" if integer else "frac"
        for dtype, sfx in ((np.float64, "f64"), (np.float32, "f32")):
            rng = np.random.default_rng(77001 + (0 if integer else 1) + (0 if sfx == "f64" else 7))
            for (Mb, Nb, br, bc, K) in ((9, 7, 4, 4, 12), (5, 6, 16, 16, 40), (4, 5, 32, 32, 128), (6, 4, 3, 5, 7)):
                blk = rng.random((Mb, Nb)) < 0.45
                blk[2 % Mb] = False                                   # an empty block row
                A = dense_fill(rng, (Mb * br, Nb * bc), integer, dtype) + (1 if integer else 0.5)
                A = (A.reshape(Mb, br, Nb, bc) * blk[:, None, :, None]).reshape(Mb * br, Nb * bc).astype(dtype)
                p, c, v = formats.bcsr_from_dense(A, br, bc)
                B = dense_fill(rng, (Nb * bc, K), integer, dtype)
                inp = dict(dims=np.array([Mb, Nb, br, bc, K], np.int32), A_pos=p, A_crd=c, A_vals=v.reshape(-1), B=B)
                out = run_ref("bspmm", inp, sfx, "default")
                cases[f"bspmm_{tag}_{sfx}_{br}x{bc}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
                if (br, bc) in ((4, 4), (3, 5)):
                    cv = dense_fill(rng, (Nb, bc), integer, dtype)
                    inp = dict(dims=np.array([Mb, Nb, br, bc], np.int32), A_pos=p, A_crd=c, A_vals=v.reshape(-1), c=cv)
                    out = run_ref("bspmv", inp, sfx, "default")
                    cases[f"bspmv_{tag}_{sfx}_{br}x{bc}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
    # ---- SpMM with a doubly compressed operand ({Sparse,Sparse}; the reference's spmmDCSRGPU test shape 102 x 103,
    #      K = 128, sparsity .3, tests-scheduling-eval.cpp:1309-1358) -- here with many absent rows as well ----------
    for integer in (True, False):
        tag = "int" if integer else "frac"
        for dtype, sfx in ((np.float64, "f64"), (np.float32, "f32")):
            rng = np.random.default_rng(55001 + (0 if integer else 1) + (0 if sfx == "f64" else 7))
            for (n, m, K, sp, keep) in ((102, 103, 128, 0.3, 1.0), (300, 90, 20, 0.1, 0.15), (64, 50, 7, 0.2, 0.5)):
                A = sparse_fill(rng, (n, m), sp, integer, dtype)
                A[rng.random(n) >= keep] = 0                          # absent rows (not stored at level 0)
                if keep < 1.0:
                    A[0] = 0
                    A[n - 1] = 0                                      # leading and trailing absent rows
                d = formats.dcsr_from_dense(A)
                B = dense_fill(rng, (m, K), integer, dtype)
                inp = dict(dims=np.array([n, m, K], np.int32), B=B, **d)
                out = run_ref("spmm_dcsr", inp, sfx, "default")
                cases[f"dcsr_spmm_{tag}_{sfx}_{n}x{K}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
    # ---- SDDMM with a dense result and D indexed (contraction, column): the reference's sddmmGPU statement
    #      (tests-scheduling-eval.cpp:1360-1418; its shape 102 x 103, contraction 128, sparsity .3) ---------------------
    for integer in (True, False):
        tag = "int" if integer else "frac"
        for dtype, sfx in ((np.float64, "f64"), (np.float32, "f32")):
            rng = np.random.default_rng(44001 + (0 if integer else 1) + (0 if sfx == "f64" else 7))
            for (n, m, J, sp) in ((102, 103, 128, 0.3), (61, 40, 7, 0.15)):
                Bm = sparse_fill(rng, (n, m), sp, integer, dtype)
                Bm[5] = 0
                p, c, v = formats.csr_from_dense(Bm)
                C = dense_fill(rng, (n, J), integer, dtype)
                D = dense_fill(rng, (J, m), integer, dtype)
                inp = dict(dims=np.array([n, m, J], np.int32), B_pos=p, B_crd=c, B_vals=v, C=C, D=D)
                out = run_ref("sddmm_dense", inp, sfx, "default")
                cases[f"densesddmm_{tag}_{sfx}_{n}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
    # ---- pack(): COO (unsorted; integer-valued cases contain duplicates) -> CSR / DCSR / CSF through the reference's
    #      insert() + pack() (src/tensor.cpp:295-463) ------------------------------------------------------------------
    for integer in (True, False):
        tag = "int" if integer else "frac"
        for dtype, sfx in ((np.float64, "f64"), (np.float32, "f32")):
            rng = np.random.default_rng(66001 + (0 if integer else 1) + (0 if sfx == "f64" else 7))
            for kind, dims, n in (("csr", (57, 43), 700), ("csr", (300, 7), 90), ("dcsr", (200, 61), 400), ("csc", (45, 71), 600),
                                  ("csf3", (23, 17, 29), 1500), ("csf3", (90, 5, 40), 300)):
                cs = [rng.integers(0, d, n).astype(np.int32) for d in dims]
                if kind != "csf3":
                    cs[0][cs[0] == 3] = 4                                     # an absent row
                if not integer:                                               # distinct coordinates: order-independent sums
                    flat = np.ravel_multi_index(cs, dims)
                    _, first = np.unique(flat, return_index=True)
                    first = rng.permutation(first)
                    cs = [c[first] for c in cs]
                v = (np.floor(rng.random(cs[0].size) * 9) + 1 if integer else rng.random(cs[0].size) + 0.25).astype(dtype)
                inp = dict(dims=np.array(dims, np.int32), vals=v, **{f"c{m}": c for m, c in enumerate(cs)})
                out = run_ref("pack_" + kind, inp, sfx, "default")
                cases[f"pack_{kind}_{tag}_{sfx}_{dims[0]}"] = dict(inp, **{"out_" + k: a for k, a in out.items()})
    only = sys.argv[1].split(",") if len(sys.argv) > 1 else [""]          # comma-separated name prefixes
    cases = {k: v for k, v in cases.items() if any(k.startswith(p) for p in only)}
    for name, arrs in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print(f"wrote {len(cases)} golden cases to {OUT}")


if __name__ == "__main__":
    main()
