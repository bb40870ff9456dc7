"""N>1 host logic on CPU: world_size-2 `gloo` processes run the partitioner + allgather path (`-m "not gpu"`).

Each rank takes its nnz-balanced shard (taco_b200.partition), computes its block of the result -- with the CPU oracle
standing in for the CUDA kernel, which is allowed in tests -- and the row blocks are all-gathered.  The gathered result
must be bit-identical to the single-process oracle on the whole operand (sharding must not change any row's
operation order)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    import synth
    from taco_b200 import partition
    try:
        # ---- CSR row sharding: iterative SpMV  x <- A x  (allgather of y between iterations) ---------------------
        w = synth.make("spmm", None, scale=12, K=8, dtype="float64")       # power-law rows: shards differ in row count
        n = w["dims"][0]
        bounds = partition.row_bounds(w["A_pos"], n, world)
        sh = partition.shard_csr(w["A_pos"], w["A_crd"], w["A_vals"], n, rank, world, bounds)
        x = synth.dense(synth.backend(None), n, 1, 99, np.float64)
        x_full = x.copy()
        for _ in range(2):
            y_local = oracle.spmv(sh["pos"], sh["crd"], sh["vals"], x)
            x = partition.allgather_rows(torch.from_numpy(y_local), bounds).numpy()
            x_full = oracle.spmv(w["A_pos"], w["A_crd"], w["A_vals"], x_full)
        ok_spmv = np.array_equal(x, x_full)
        # ---- row-sharded SpMM: C blocks gathered --------------------------------------------------------------
        B = w["B"].reshape(n, 8)
        c_local = oracle.spmm(sh["pos"], sh["crd"], sh["vals"], B)
        C = partition.allgather_rows(torch.from_numpy(c_local), bounds, row_len=8).numpy().reshape(n, 8)
        ok_spmm = np.array_equal(C, oracle.spmm(w["A_pos"], w["A_crd"], w["A_vals"], B))
        # ---- CSF mode-0 slice sharding: MTTKRP ---------------------------------------------------------------
        t = synth.make("mttkrp", None, I=3000, K=200, L=150, nnz=40_000, R=8, dtype="float64")
        I, K, L, R = t["dims"]
        st = partition.shard_csf3(t, rank, world)
        a_local = oracle.mttkrp(st, t["C"].reshape(K, R), t["D"].reshape(L, R), I)   # rows outside the shard are 0
        a_sum = torch.from_numpy(a_local.copy())
        dist.all_reduce(a_sum)                      # disjoint row ownership: the sum is a concatenation
        ok_mttkrp = np.array_equal(a_sum.numpy(), oracle.mttkrp(t, t["C"].reshape(K, R), t["D"].reshape(L, R), I))
        nnz_local = int(st["B_vals"].shape[0])
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([nnz_local]))
        total = sum(int(s) for s in sizes)
        balanced = max(int(s) for s in sizes) <= total / world + 64
        covered = total == int(t["B_vals"].shape[0])
        results[rank] = (ok_spmv, ok_spmm, ok_mttkrp, balanced, covered, int(sh["row_end"] - sh["row_begin"]))
    finally:
        dist.destroy_process_group()


def test_partitioner_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert len(results) == world
        rows = []
        for r in range(world):
            ok_spmv, ok_spmm, ok_mttkrp, balanced, covered, nrows = results[r]
            assert ok_spmv, "sharded iterative SpMV + allgather differs from the single-process oracle"
            assert ok_spmm, "row-sharded SpMM differs"
            assert ok_mttkrp, "slice-sharded MTTKRP differs"
            assert balanced and covered
            rows.append(nrows)
        assert sum(rows) == 1 << 12


def test_shard_rebase_is_consistent():
    import synth
    from taco_b200 import partition
    w = synth.make("spmv", None, n=1000, deg=7)
    for world in (1, 3, 8):
        b = partition.row_bounds(w["A_pos"], 1000, world)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        tot = 0
        for r in range(world):
            sh = partition.shard_csr(w["A_pos"], w["A_crd"], w["A_vals"], 1000, r, world, b)
            assert sh["pos"][0] == 0 and sh["pos"][-1] == sh["crd"].shape[0] == sh["vals"].shape[0]
            tot += int(sh["crd"].shape[0])
        assert tot == 7000


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_csf_shards_with_rebased_rows_are_self_contained(world):
    """shard_csf3(rebase_rows=True): every shard is an MTTKRP over its OWN rows of the result (row ids rebased, contiguous row
    ranges that tile [0, I)); the concatenated blocks equal the single-process oracle bit for bit"""
    import oracle
    import synth
    from taco_b200 import partition
    t = synth.make("mttkrp", None, I=3000, K=200, L=150, nnz=40_000, R=8, dtype="float64")
    I, K, L, R = t["dims"]
    full = oracle.mttkrp(t, t["C"].reshape(K, R), t["D"].reshape(L, R), I)
    blocks, nxt = [], 0
    for r in range(world):
        st = partition.shard_csf3(t, r, world, rebase_rows=True, dim0=I)
        assert st["row_begin"] == nxt and st["row_end"] >= st["row_begin"]
        nxt = st["row_end"]
        blocks.append(oracle.mttkrp(st, t["C"].reshape(K, R), t["D"].reshape(L, R), st["row_end"] - st["row_begin"]))
    assert nxt == I
    assert np.array_equal(np.concatenate(blocks), full)
