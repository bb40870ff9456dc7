"""Bring-up aid for the tcgen05 blocked SpMM: single-block probes whose wrong entries say which operand chunk / TMEM column
is misplaced.  Run on a GPU box: python tools/bcsr_debug.py [br]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import gpu_util as G  # noqa: E402


def run(A, B, br, bc, K):
    pos = np.array([0, 1], np.int32)
    crd = np.array([0], np.int32)
    C = G.run("bspmm", dict(dims=[1, 1, br, bc, K], A_pos=pos, A_crd=crd, A_vals=A.reshape(-1).astype(np.float32),
                            B=B.reshape(-1).astype(np.float32)))
    return C.reshape(br, K)


def main():
    br = bc = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    K = 128
    rng = np.random.default_rng(0)
    # 1. exact-in-tf32 integers: only the hi*hi term is non-zero
    A = rng.integers(-3, 4, (br, bc)).astype(np.float32)
    B = rng.integers(-3, 4, (bc, K)).astype(np.float32)
    C = run(A, B, br, bc, K)
    want = A @ B
    bad = np.argwhere(C != want)
    print(f"[int] wrong {len(bad)} / {C.size}; rows {sorted(set(bad[:, 0]))[:40]} cols {sorted(set(bad[:, 1]))[:20]}")
    # 2. unit probes: A = e(i2, j2), B = row index coded -> C[i2, :] must be B[j2, :]
    for (i2, j2) in [(0, 0), (1, 0), (0, 1), (8, 0), (0, 4), (0, 8), (br - 1, bc - 1), (br // 2, 5)]:
        A = np.zeros((br, bc), np.float32)
        A[i2, j2] = 1.0
        B = (np.arange(bc)[:, None] * 1000 + np.arange(K)[None, :] + 1).astype(np.float32)
        C = run(A, B, br, bc, K)
        nz = np.argwhere(C != 0)
        rows = sorted(set(nz[:, 0]))
        got = C[rows[0], :4] if rows else None
        print(f"[unit A({i2},{j2})] non-zero rows {rows[:8]} first values {got} want row {i2} values {B[j2, :4]}")
    # 3. lo terms: values with low mantissa bits
    A = (rng.random((br, bc)) + 1e-4).astype(np.float32)
    B = (rng.random((bc, K)) + 1e-4).astype(np.float32)
    C = run(A, B, br, bc, K)
    want = A.astype(np.float64) @ B.astype(np.float64)
    print(f"[frac] max rel err {np.abs(C - want).max() / np.abs(want).max():.3e}")
    Ahi = (A.view(np.uint32) & 0xFFFFE000).view(np.float32)
    Bhi = (B.view(np.uint32) & 0xFFFFE000).view(np.float32)
    for name, w in [("hi*hi", Ahi.astype(np.float64) @ Bhi), ("hi*hi+lo*hi", A.astype(np.float64) @ Bhi),
                    ("hi*hi+hi*lo", Ahi.astype(np.float64) @ B)]:
        print(f"   vs {name}: {np.abs(C - w).max() / np.abs(want).max():.3e}")


if __name__ == "__main__":
    main()
