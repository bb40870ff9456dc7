"""Multi-GPU partitioner (new: the reference has no multi-device support at all, SURVEY.md sections 2 #25, 8(e)).

One process per GPU.  The sparse operand is split into contiguous, nnz-balanced pieces of its OUTER level:
  * CSR  (SpMV / SpMM / SDDMM / SpAdd / SpGEMM): row ranges; pos is rebased to 0, crd/vals are slices,
  * CSF  (MTTKRP / TTV / TTM): mode-0 slice ranges balanced by leaf count; every level is rebased,
with the split points found by the same "first position >= target" search the reference uses to hand nnz ranges to
thread blocks (taco_binarySearchBeforeBlock, /root/reference/src/codegen/codegen_cuda.cpp:110-125) -- here via
taco_b200_partition_pos, applied across GPUs.  Dense operands (x, B, the factor matrices) are replicated.
Each rank then produces a disjoint row block of the dense result with NO data-path collective; the only exchange
is the allgather of those row blocks when the next iteration needs the full dense operand (e.g. y -> x in an
iterative SpMV, the updated factor in an ALS sweep): `allgather_rows`, NCCL over NVLink on GPUs, gloo in CPU tests.
"""
import numpy as np

from .tensor import partition_pos

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


def _rebase(pos_slice):
    return pos_slice - pos_slice[0]


def row_bounds(pos, n_rows, world):
    """bounds[g]..bounds[g+1] = rows of shard g (balanced by nnz)"""
    return partition_pos(pos, n_rows, world)


def shard_csr(pos, crd, vals, n_rows, rank, world, bounds=None):
    """Row shard `rank` of a CSR operand -> dict(pos, crd, vals, row_begin, row_end)."""
    if bounds is None:
        bounds = row_bounds(pos, n_rows, world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    p = pos[r0:r1 + 1]
    lo, hi = int(p[0]), int(p[-1])
    lp = _rebase(p)
    lp = lp.contiguous() if _is_torch(lp) else np.ascontiguousarray(lp)
    return dict(pos=lp, crd=crd[lo:hi], vals=vals[lo:hi], row_begin=r0, row_end=r1)


def shard_csf3(t, rank, world, rebase_rows=False, dim0=None):
    """Mode-0 slice shard of a CSF tensor (dict B1_pos.. as produced by formats.coo_to_csf3 / synth.csf3_uniform).

    rebase_rows=True makes the shard a self-contained sub-problem over ITS OWN rows of the dense result (what a rank
    of a multi-GPU MTTKRP runs): the shard owns the result rows [row_begin, row_end) -- from its first slice's row id
    (0 for rank 0) up to the next shard's first row id (`dim0` for the last rank) -- and B1_crd is rebased to that
    range, so the kernel writes a (row_end - row_begin) x R block that concatenates with the other ranks' blocks."""
    ns = int(t["B1_crd"].shape[0])
    # leaf offset at the start of every slice: B3_pos[B2_pos[s]]
    idx = t["B2_pos"].long() if _is_torch(t["B2_pos"]) else t["B2_pos"].astype(np.int64)
    slice_leaf = t["B3_pos"][idx]
    slice_leaf = slice_leaf.contiguous() if _is_torch(slice_leaf) else np.ascontiguousarray(slice_leaf)
    bounds = partition_pos(slice_leaf, ns, world)
    s0, s1 = int(bounds[rank]), int(bounds[rank + 1])
    p2 = t["B2_pos"][s0:s1 + 1]
    f0, f1 = int(p2[0]), int(p2[-1])
    p3 = t["B3_pos"][f0:f1 + 1]
    l0, l1 = int(p3[0]), int(p3[-1])
    mk = (lambda a: a.contiguous()) if _is_torch(t["B2_pos"]) else np.ascontiguousarray
    b1 = np.array([0, s1 - s0], dtype=np.int32)
    if _is_torch(t["B2_pos"]):
        b1 = torch.as_tensor(b1, device=t["B2_pos"].device)
    c1 = t["B1_crd"][s0:s1]
    out = dict(B1_pos=b1, B2_pos=mk(_rebase(p2)), B2_crd=mk(t["B2_crd"][f0:f1]),
               B3_pos=mk(_rebase(p3)), B3_crd=mk(t["B3_crd"][l0:l1]), B_vals=mk(t["B_vals"][l0:l1]),
               slice_begin=s0, slice_end=s1)
    if rebase_rows:
        assert dim0 is not None, "rebase_rows needs the mode-0 dimension"
        row_begin = 0 if rank == 0 else (int(t["B1_crd"][s0]) if s0 < ns else int(dim0))
        row_end = int(t["B1_crd"][s1]) if s1 < ns else int(dim0)
        c1 = c1 - row_begin
        out.update(row_begin=row_begin, row_end=row_end)
    out["B1_crd"] = mk(c1)
    return out


def allgather_rows(local_rows, bounds, row_len=1, group=None):
    """All-gather contiguous row blocks of a dense result: rank g contributes rows bounds[g]..bounds[g+1] (each
    `row_len` values).  Returns the full (bounds[-1] * row_len) tensor on every rank.  Uneven blocks are padded to
    the largest one for the collective (one NCCL allgather; NVSwitch makes it bandwidth- not link-bound)."""
    world = dist.get_world_size(group)
    sizes = [int(bounds[g + 1] - bounds[g]) * row_len for g in range(world)]
    mx = max(sizes)
    buf = torch.zeros(mx, dtype=local_rows.dtype, device=local_rows.device)
    buf[: local_rows.numel()] = local_rows.reshape(-1)
    out = torch.empty(world * mx, dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return torch.cat([out[g * mx: g * mx + sizes[g]] for g in range(world)])
