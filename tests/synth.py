"""Deterministic synthetic operands for the BASELINE.json configs.

The reference's own `-g` fills are time-seeded (/root/reference/include/taco/util/fill.h:102-144), so the
workloads are defined here.  Everything is derived from a counter-based hash (splitmix64 of the element
index), so the SAME arrays come out of numpy on the host and torch on the GPU: small cases are generated on
the host for the oracle, full-size cases directly in HBM.

Configs ("workloads"):
  C1 spmv   fp64  n x n CSR, exactly `deg` nnz per row, columns uniform (one per stratum of width n/deg => sorted,
                  unique, marginally uniform);  x uniform
  C2 spmm   fp32  R-MAT power-law (a,b,c,d = .57,.19,.19,.05; Graph500 parameters) 2^s x 2^s, edge factor 16,
                  vertex labels scrambled by a bijective hash, duplicates removed;  B dense n x K
  C3 sddmm  fp32  n x n CSR `deg` per row (as C1);  C, D dense n x K
  C4 mttkrp fp64  order-3 I x K x L, coordinates uniform, sorted, duplicates removed, CSF;  C, D dense x R
  C5 spadd / spgemm fp64  two independent C1-style matrices
  (f)1 bspmm fp32  blocked CSR {Dense,Compressed,Dense,Dense}: C1-style block structure, dense 32 x 32 blocks
Values are k/1024 with k uniform in [1, 1024] (well-conditioned sums, never zero).
"""
import numpy as np

try:  # torch is optional on the host path
    import torch
except Exception:  # pragma: no cover
    torch = None

_M64 = (1 << 64) - 1


def _s64(c):
    """python int (u64 constant) -> the same bits as signed int64"""
    c &= _M64
    return c - (1 << 64) if c >= (1 << 63) else c


_GOLD = _s64(0x9E3779B97F4A7C15)
_MIX1 = _s64(0xBF58476D1CE4E5B9)
_MIX2 = _s64(0x94D049BB133111EB)


class _NP:
    name = "numpy"

    def arange(self, n):
        return np.arange(n, dtype=np.int64)

    def lsr(self, z, s):
        return z if s == 0 else (z >> s) & np.int64((1 << (64 - s)) - 1)

    def mul(self, a, c):
        with np.errstate(over="ignore"):
            return a * np.int64(c)

    def add(self, a, c):
        with np.errstate(over="ignore"):
            return a + np.int64(c)

    def f64(self, a):
        return a.astype(np.float64)

    def i64(self, a):
        return a.astype(np.int64)

    def i32(self, a):
        return a.astype(np.int32)

    def cast(self, a, dtype):
        return a.astype(np.dtype(dtype))

    def sort(self, a):
        return np.sort(a, kind="stable")

    def argsort(self, a):
        return np.argsort(a, kind="stable")

    def cumsum(self, a):
        return np.cumsum(a)

    def nonzero(self, m):
        return np.flatnonzero(m)

    def ones_bool(self, n):
        return np.ones(n, dtype=bool)

    def zeros(self, n, dtype=np.int64):
        return np.zeros(n, dtype=dtype)

    def cat(self, xs):
        return np.concatenate(xs)

    def bincount(self, a, n):
        return np.bincount(a, minlength=n).astype(np.int64)

    def scalar(self, v):
        return np.array([v], dtype=np.int64)


class _TH:
    name = "torch"

    def __init__(self, device):
        self.device = torch.device(device)

    def arange(self, n):
        return torch.arange(n, dtype=torch.int64, device=self.device)

    def lsr(self, z, s):
        return z if s == 0 else (z >> s) & ((1 << (64 - s)) - 1)

    def mul(self, a, c):
        return a * c

    def add(self, a, c):
        return a + c

    def f64(self, a):
        return a.to(torch.float64)

    def i64(self, a):
        return a.to(torch.int64)

    def i32(self, a):
        return a.to(torch.int32)

    def cast(self, a, dtype):
        return a.to({np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)])

    def sort(self, a):
        return torch.sort(a, stable=True).values

    def argsort(self, a):
        return torch.sort(a, stable=True).indices

    def cumsum(self, a):
        return torch.cumsum(a, 0)

    def nonzero(self, m):
        return torch.nonzero(m).reshape(-1)

    def ones_bool(self, n):
        return torch.ones(n, dtype=torch.bool, device=self.device)

    def zeros(self, n, dtype=None):
        return torch.zeros(n, dtype=torch.int64, device=self.device)

    def cat(self, xs):
        return torch.cat(xs)

    def bincount(self, a, n):
        return torch.bincount(a, minlength=n).to(torch.int64)

    def scalar(self, v):
        return torch.tensor([v], dtype=torch.int64, device=self.device)


def backend(device=None):
    """device None/'cpu-numpy' -> numpy arrays; otherwise a torch device string ('cuda', 'cuda:0', 'cpu')."""
    if device is None or device == "numpy":
        return _NP()
    if torch is None:
        raise RuntimeError("torch not importable")
    return _TH(device)


def hash64(xp, idx, seed):
    """splitmix64 finaliser of (idx * golden + seed); returns int64 with uniformly random bits."""
    z = xp.add(xp.mul(idx, _GOLD), _s64(seed * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D))
    z = xp.mul(z ^ xp.lsr(z, 30), _MIX1)
    z = xp.mul(z ^ xp.lsr(z, 27), _MIX2)
    return z ^ xp.lsr(z, 31)


def uniform_int(xp, idx, seed, n):
    """uniform integer in [0, n) from 53 hashed bits"""
    h = xp.lsr(hash64(xp, idx, seed), 11)          # 53 bits, non-negative
    return xp.i64(xp.f64(h) * (float(n) / float(1 << 53)))


def values(xp, idx, seed, dtype):
    """k/1024, k in [1, 1024]"""
    k = xp.lsr(hash64(xp, idx, seed), 54) + 1      # 10 bits -> [1, 1024]
    return xp.cast(xp.f64(k) * (1.0 / 1024.0), dtype)


def dense(xp, rows, cols, seed, dtype):
    """row-major dense operand (flattened length rows*cols)"""
    return values(xp, xp.arange(rows * cols), seed, dtype)


def csr_fixed_degree(xp, n_rows, n_cols, deg, seed, dtype):
    """CSR with exactly `deg` entries per row; entry t of a row falls uniformly in column stratum t."""
    assert n_cols >= deg
    nnz = n_rows * deg
    e = xp.arange(nnz)
    t = e % deg
    lo = (t * n_cols) // deg
    hi = ((t + 1) * n_cols) // deg
    crd = lo + uniform_int(xp, e, seed, 1 << 30) % (hi - lo)
    pos = xp.arange(n_rows + 1) * deg
    return xp.i32(pos), xp.i32(crd), values(xp, e, seed + 1, dtype)


def _scramble(xp, v, bits, seed):
    """bijection on [0, 2^bits): odd multiply + xorshift rounds (Graph500-style vertex relabelling)"""
    mask = (1 << bits) - 1
    for r in range(3):
        c = (_s64(hash_const(seed + r)) | 1)
        v = xp.mul(v, c) & mask
        v = v ^ (v >> max(1, bits // 2))
    return v


def hash_const(seed):
    z = (seed * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) & _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def _rmat_keys(xp, scale, m, seed, abcd):
    n = 1 << scale
    e = xp.arange(m)
    a, b, c, _ = abcd
    ta, tb, tc = int(a * 65536), int((a + b) * 65536), int((a + b + c) * 65536)
    row = xp.zeros(m)
    col = xp.zeros(m)
    for lvl in range(scale):
        if lvl % 4 == 0:
            h = hash64(xp, e, seed + 101 * (lvl // 4))
        u = xp.lsr(h, 16 * (lvl % 4)) & 0xFFFF
        rbit = xp.i64(u >= tb)                      # quadrants c, d -> lower half (row bit 1)
        cbit = xp.i64(((u >= ta) & (u < tb)) | (u >= tc))   # quadrants b, d -> right half
        row = row * 2 + rbit
        col = col * 2 + cbit
    row = _scramble(xp, row, scale, seed + 7)
    col = _scramble(xp, col, scale, seed + 13)
    key = xp.sort(row * n + col)
    keep = xp.ones_bool(m)
    keep[1:] = key[1:] != key[:-1]
    return key[keep]


def csr_rmat(xp, scale, edge_factor, seed, dtype, abcd=(0.57, 0.19, 0.19, 0.05)):
    """R-MAT power-law matrix 2^scale x 2^scale with exactly min(edge_factor * 2^scale, 4^scale / 2) unique
    entries: edges are oversampled (1.35x, growing until enough distinct ones exist), de-duplicated, and a
    deterministic pseudo-random subset is dropped to hit the target count exactly."""
    n = 1 << scale
    target = min(n * edge_factor, (n * n) // 2)
    for oversample in (1.35, 2.0, 3.0, 5.0, 8.0, 16.0, 64.0):
        key = _rmat_keys(xp, scale, int(target * oversample), seed, abcd)
        if int(key.shape[0]) >= target:
            break
    uniq = int(key.shape[0])
    assert uniq >= target, f"R-MAT draw produced {uniq} distinct edges < {target}"
    if uniq > target:
        pri = xp.lsr(hash64(xp, key, seed + 29), 1)
        thresh = xp.sort(pri)[target - 1]
        key = key[pri <= thresh][:target]
    row = key // n
    crd = key % n
    pos = xp.cat([xp.scalar(0), xp.cumsum(xp.bincount(row, n))])
    return xp.i32(pos), xp.i32(crd), values(xp, xp.arange(int(key.shape[0])), seed + 1, dtype)


def _csf3_from_coords(xp, ik, l, K, nnz, seed, dtype):
    """sort (ik = i*K+k, l) lexicographically with two stable sorts (the full key would not fit 63 bits at
    10M x 1M x 1M), drop duplicates, build the three CSF levels"""
    o = xp.argsort(l)
    l, ik = l[o], ik[o]
    o = xp.argsort(ik)
    l, ik = l[o], ik[o]
    del o
    keep = xp.ones_bool(nnz)
    keep[1:] = (ik[1:] != ik[:-1]) | (l[1:] != l[:-1])
    ik, l = ik[keep], l[keep]
    nnz = int(ik.shape[0])
    new_fib = xp.ones_bool(nnz)
    new_fib[1:] = ik[1:] != ik[:-1]
    fib_start = xp.nonzero(new_fib)
    fib_ik = ik[fib_start]
    fib_i = fib_ik // K
    nf = int(fib_start.shape[0])
    new_slice = xp.ones_bool(nf)
    new_slice[1:] = fib_i[1:] != fib_i[:-1]
    slice_start = xp.nonzero(new_slice)
    return dict(
        B1_pos=xp.i32(xp.cat([xp.scalar(0), xp.scalar(int(slice_start.shape[0]))])),
        B1_crd=xp.i32(fib_i[slice_start]),
        B2_pos=xp.i32(xp.cat([slice_start, xp.scalar(nf)])),
        B2_crd=xp.i32(fib_ik % K),
        B3_pos=xp.i32(xp.cat([fib_start, xp.scalar(nnz)])),
        B3_crd=xp.i32(l),
        B_vals=values(xp, xp.arange(nnz), seed + 3, dtype),
    )


def csf3_uniform(xp, I, K, L, nnz, seed, dtype):
    """order-3 CSF {Compressed x3} (mode order 0,1,2) with ~nnz uniform coordinates (duplicates removed)."""
    e = xp.arange(nnz)
    i = uniform_int(xp, e, seed, I)
    k = uniform_int(xp, e, seed + 1, K)
    l = uniform_int(xp, e, seed + 2, L)
    assert I * K < (1 << 62)
    ik = i * K + k
    del i, k, e
    return _csf3_from_coords(xp, ik, l, K, nnz, seed, dtype)


def csf3_fibers(xp, I, K, L, nnz, nfib, seed, dtype):
    """order-3 CSF whose (i,k) fibers hold about nnz / nfib leaves each (real tensors have fibers longer than one
    leaf; csf3_uniform at the C4 shape has nfib ~ nnz): every nonzero draws one of `nfib` fibers, the fiber's (i,k) cell
    is a hash of the fiber id, l is uniform."""
    e = xp.arange(nnz)
    f = uniform_int(xp, e, seed, nfib)
    i = uniform_int(xp, f, seed + 1, I)
    k = uniform_int(xp, f, seed + 4, K)
    l = uniform_int(xp, e, seed + 2, L)
    assert I * K < (1 << 62)
    ik = i * K + k
    del i, k, e, f
    return _csf3_from_coords(xp, ik, l, K, nnz, seed, dtype)


# ---- the named BASELINE workloads ---------------------------------------------------------------------------
SEED0 = 0x7AC00000

FULL = {
    "spmv": dict(n=1_000_000, deg=10, dtype="float64"),
    "spmm": dict(scale=22, edge_factor=16, K=128, dtype="float32"),
    "sddmm": dict(n=2_000_000, deg=20, K=64, dtype="float32"),
    "mttkrp": dict(I=10_000_000, K=1_000_000, L=1_000_000, nnz=200_000_000, R=32, dtype="float64"),
    # a second MTTKRP tensor with realistic fiber lengths (about 12 leaves per (i,k) fiber, 4 fibers per slice): the C4 tensor
    # above has one leaf per fiber, which hides what a per-fiber workspace saves
    "mttkrp_fibers": dict(I=2_000_000, K=1_000_000, L=1_000_000, nnz=100_000_000, nfib=8_000_000, R=32, dtype="float64"),
    "spadd": dict(n=1_000_000, deg=10, dtype="float64"),
    "spgemm": dict(n=1_000_000, deg=10, dtype="float64"),
    # TTV / TTM: dense results, so the (i,j) plane is kept small enough to hold A in HBM
    "ttv": dict(I=8192, K=8192, L=1_000_000, nnz=100_000_000, dtype="float64"),
    "ttm": dict(I=1024, K=1024, L=1_000_000, nnz=50_000_000, R=32, dtype="float64"),
    # pack(): C1's shape as unsorted coordinates, uniform (a few duplicates), -> CSR
    "pack": dict(n=1_000_000, nnz=10_000_000, dtype="float64"),
    # blocked SpMM: 1Mi x 1Mi in 32 x 32 blocks, 16 stored blocks per block row, K = 128
    "bspmm": dict(Mb=32768, deg=16, br=32, bc=32, K=128, dtype="float32"),
    # blocked SpMV, the reference's `bspmv` statement (tests-expr_storage.cpp:939-960), on the same block structure in fp64
    "bspmv": dict(Mb=32768, deg=16, br=32, bc=32, dtype="float64"),
}


def make(workload, device=None, **over):
    """Build the operands of a named workload (FULL sizes unless overridden). Returns a dict of arrays + 'dims'."""
    xp = backend(device)
    p = dict(FULL[workload])
    p.update(over)
    dt = np.dtype(p["dtype"])
    if workload == "spmv":
        pos, crd, vals = csr_fixed_degree(xp, p["n"], p["n"], p["deg"], SEED0 + 1, dt)
        return dict(dims=(p["n"], p["n"]), A_pos=pos, A_crd=crd, A_vals=vals, x=dense(xp, p["n"], 1, SEED0 + 3, dt))
    if workload == "spmm":
        n = 1 << p["scale"]
        pos, crd, vals = csr_rmat(xp, p["scale"], p["edge_factor"], SEED0 + 2, dt)
        return dict(dims=(n, n, p["K"]), A_pos=pos, A_crd=crd, A_vals=vals, B=dense(xp, n, p["K"], SEED0 + 4, dt))
    if workload == "sddmm":
        pos, crd, vals = csr_fixed_degree(xp, p["n"], p["n"], p["deg"], SEED0 + 5, dt)
        return dict(dims=(p["n"], p["n"], p["K"]), B_pos=pos, B_crd=crd, B_vals=vals,
                    C=dense(xp, p["n"], p["K"], SEED0 + 7, dt), D=dense(xp, p["n"], p["K"], SEED0 + 8, dt))
    if workload == "mttkrp":
        t = csf3_uniform(xp, p["I"], p["K"], p["L"], p["nnz"], SEED0 + 9, dt)
        t.update(dims=(p["I"], p["K"], p["L"], p["R"]), C=dense(xp, p["K"], p["R"], SEED0 + 14, dt),
                 D=dense(xp, p["L"], p["R"], SEED0 + 15, dt))
        return t
    if workload == "mttkrp_fibers":
        t = csf3_fibers(xp, p["I"], p["K"], p["L"], p["nnz"], p["nfib"], SEED0 + 50, dt)
        t.update(dims=(p["I"], p["K"], p["L"], p["R"]), C=dense(xp, p["K"], p["R"], SEED0 + 54, dt),
                 D=dense(xp, p["L"], p["R"], SEED0 + 55, dt))
        return t
    if workload == "ttv":
        t = csf3_uniform(xp, p["I"], p["K"], p["L"], p["nnz"], SEED0 + 24, dt)
        t.update(dims=(p["I"], p["K"], p["L"]), c=dense(xp, p["L"], 1, SEED0 + 28, dt))
        return t
    if workload == "ttm":
        t = csf3_uniform(xp, p["I"], p["K"], p["L"], p["nnz"], SEED0 + 30, dt)
        t.update(dims=(p["I"], p["K"], p["L"], p["R"]), C=dense(xp, p["L"], p["R"], SEED0 + 34, dt))
        return t
    if workload == "pack":
        e = xp.arange(p["nnz"])
        return dict(dims=(p["n"], p["n"]), c0=xp.i32(uniform_int(xp, e, SEED0 + 40, p["n"])),
                    c1=xp.i32(uniform_int(xp, e, SEED0 + 41, p["n"])), vals=values(xp, e, SEED0 + 42, dt))
    if workload in ("spadd", "spgemm"):
        ap, ac, av = csr_fixed_degree(xp, p["n"], p["n"], p["deg"], SEED0 + 16, dt)
        bp, bc, bv = csr_fixed_degree(xp, p["n"], p["n"], p["deg"], SEED0 + 18, dt)
        dims = (p["n"], p["n"]) if workload == "spadd" else (p["n"], p["n"], p["n"])
        return dict(dims=dims, A_pos=ap, A_crd=ac, A_vals=av, B_pos=bp, B_crd=bc, B_vals=bv)
    if workload == "bspmv":
        Mb, br, bc = p["Mb"], p["br"], p["bc"]
        pos, crd, _ = csr_fixed_degree(xp, Mb, Mb, p["deg"], SEED0 + 20, dt)
        nnzb = Mb * p["deg"]
        return dict(dims=(Mb, Mb, br, bc), A_pos=pos, A_crd=crd, A_vals=values(xp, xp.arange(nnzb * br * bc), SEED0 + 21, dt),
                    c=dense(xp, Mb * bc, 1, SEED0 + 23, dt))
    if workload == "bspmm":
        Mb, br, bc = p["Mb"], p["br"], p["bc"]
        pos, crd, _ = csr_fixed_degree(xp, Mb, Mb, p["deg"], SEED0 + 20, dt)
        nnzb = Mb * p["deg"]
        return dict(dims=(Mb, Mb, br, bc, p["K"]), A_pos=pos, A_crd=crd,
                    A_vals=values(xp, xp.arange(nnzb * br * bc), SEED0 + 21, dt),
                    B=dense(xp, Mb * bc, p["K"], SEED0 + 22, dt))
    raise KeyError(workload)
