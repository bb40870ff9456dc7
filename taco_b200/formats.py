"""Host-side (numpy) builders for the level formats the hot path consumes.

Level semantics follow the reference exactly:
  * compressed level l: pos = int32[parent_size + 1], crd = int32[pos[parent_size]]
    (/root/reference/src/lower/mode_format_compressed.cpp:80-105)
  * dense level: position = parent_pos * dim + coord (/root/reference/src/lower/mode_format_dense.cpp:50-55)
  * CSR  = {Dense, Compressed}; CSF(3) = {Compressed, Compressed, Compressed}
  * BCSR = {Dense, Compressed, Dense, Dense} over (block row, block column, row in block, column in block), the
    blocked format of the reference's `bspmv` test (/root/reference/test/tests-expr_storage.cpp:939-960):
    vals = [number of stored blocks][br][bc]
"""
import numpy as np


def coo_to_csr(n_rows, rows, cols, vals):
    """Sorted, de-duplicated (last value wins is NOT applied: duplicates are summed) COO -> CSR."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals)
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    if rows.size:
        keep = np.ones(rows.size, dtype=bool)
        keep[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
        if not keep.all():
            seg = np.cumsum(keep) - 1
            summed = np.zeros(int(seg[-1]) + 1, dtype=vals.dtype)
            np.add.at(summed, seg, vals)
            rows, cols, vals = rows[keep], cols[keep], summed
    pos = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(pos, rows + 1, 1)
    pos = np.cumsum(pos)
    return pos.astype(np.int32), cols.astype(np.int32), np.ascontiguousarray(vals)


def csr_from_dense(dense):
    dense = np.asarray(dense)
    r, c = np.nonzero(dense)
    return coo_to_csr(dense.shape[0], r, c, dense[r, c])


def csr_to_dense(n_rows, n_cols, pos, crd, vals):
    out = np.zeros((n_rows, n_cols), dtype=vals.dtype)
    rows = np.repeat(np.arange(n_rows), np.diff(pos))
    np.add.at(out, (rows, crd), vals)
    return out


def coo_to_csf3(i, k, l, vals):
    """Unique COO coordinates (any order) -> CSF {Compressed,Compressed,Compressed}, mode order 0,1,2.

    Returns dict with B1_pos, B1_crd, B2_pos, B2_crd, B3_pos, B3_crd, B_vals.
    """
    i = np.asarray(i, dtype=np.int64)
    k = np.asarray(k, dtype=np.int64)
    l = np.asarray(l, dtype=np.int64)
    vals = np.asarray(vals)
    order = np.lexsort((l, k, i))
    i, k, l, vals = i[order], k[order], l[order], vals[order]
    nnz = i.size
    if nnz == 0:
        z = np.zeros(0, dtype=np.int32)
        return dict(B1_pos=np.zeros(2, dtype=np.int32), B1_crd=z, B2_pos=np.zeros(1, dtype=np.int32), B2_crd=z,
                    B3_pos=np.zeros(1, dtype=np.int32), B3_crd=z, B_vals=vals)
    new_i = np.ones(nnz, dtype=bool)
    new_i[1:] = i[1:] != i[:-1]
    new_ik = new_i.copy()
    new_ik[1:] |= k[1:] != k[:-1]
    fib_start = np.flatnonzero(new_ik)            # leaf index where each (i,k) fiber starts
    slice_start_fib = np.flatnonzero(new_i[fib_start])  # fiber index where each i slice starts
    b3_pos = np.append(fib_start, nnz)
    b2_crd = k[fib_start]
    b2_pos = np.append(slice_start_fib, fib_start.size)
    b1_crd = i[fib_start[slice_start_fib]]
    b1_pos = np.array([0, b1_crd.size])
    return dict(B1_pos=b1_pos.astype(np.int32), B1_crd=b1_crd.astype(np.int32),
                B2_pos=b2_pos.astype(np.int32), B2_crd=b2_crd.astype(np.int32),
                B3_pos=b3_pos.astype(np.int32), B3_crd=l.astype(np.int32), B_vals=np.ascontiguousarray(vals))


def csf3_to_coo(t):
    """Inverse of coo_to_csf3 (for dense cross-checks on small tensors)."""
    n_fib = t["B2_crd"].size
    fib_slice = np.repeat(np.arange(t["B1_crd"].size), np.diff(t["B2_pos"]))
    leaf_fib = np.repeat(np.arange(n_fib), np.diff(t["B3_pos"]))
    i = t["B1_crd"][fib_slice][leaf_fib]
    k = t["B2_crd"][leaf_fib]
    return i, k, t["B3_crd"], t["B_vals"]


def bcsr_from_dense(dense, br, bc):
    """dense (Mb*br, Nb*bc) -> BCSR pos[Mb+1], crd[nnzb], vals[nnzb, br, bc]; a block is stored iff it has a nonzero."""
    dense = np.asarray(dense)
    Mb, Nb = dense.shape[0] // br, dense.shape[1] // bc
    assert dense.shape == (Mb * br, Nb * bc)
    blocks = dense.reshape(Mb, br, Nb, bc).transpose(0, 2, 1, 3)         # (Mb, Nb, br, bc)
    keep = (blocks != 0).any(axis=(2, 3))
    bi, bj = np.nonzero(keep)
    pos = np.zeros(Mb + 1, dtype=np.int64)
    np.add.at(pos, bi + 1, 1)
    return np.cumsum(pos).astype(np.int32), bj.astype(np.int32), np.ascontiguousarray(blocks[bi, bj])


def bcsr_to_dense(Mb, Nb, br, bc, pos, crd, vals):
    out = np.zeros((Mb, Nb, br, bc), dtype=np.asarray(vals).dtype)
    rows = np.repeat(np.arange(Mb), np.diff(pos))
    out[rows, crd] = np.asarray(vals).reshape(-1, br, bc)
    return out.transpose(0, 2, 1, 3).reshape(Mb * br, Nb * bc)


def dcsr_from_dense(dense):
    """dense (n, m) -> doubly compressed rows {Sparse,Sparse}: A1_pos[2], A1_crd[stored rows], A2_pos[stored rows+1],
    A2_crd[nnz], A_vals[nnz]; a row is stored iff it has a nonzero (what taco's pack() produces)."""
    dense = np.asarray(dense)
    i, j = np.nonzero(dense)
    rows, counts = np.unique(i, return_counts=True)
    pos2 = np.zeros(rows.size + 1, dtype=np.int64)
    np.cumsum(counts, out=pos2[1:])
    return dict(A1_pos=np.array([0, rows.size], np.int32), A1_crd=rows.astype(np.int32), A2_pos=pos2.astype(np.int32),
                A2_crd=j.astype(np.int32), A_vals=np.ascontiguousarray(dense[i, j]))


def dcsr_to_csr(n_rows, d):
    """expand {Sparse,Sparse} level arrays to a CSR pos array over all n_rows rows (crd / vals are shared)"""
    pos = np.zeros(n_rows + 1, dtype=np.int64)
    pos[np.asarray(d["A1_crd"], dtype=np.int64) + 1] = np.diff(np.asarray(d["A2_pos"], dtype=np.int64))
    return np.cumsum(pos).astype(np.int32)
