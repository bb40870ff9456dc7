"""Multi-GPU path on real GPUs (`-m gpu`): the CUDA kernels run on the shards taco_b200.partition produces.

  * one GPU is enough for the first group: every shard of a 2- / 3- / 8-way split (device-resident VIEWS of the
    full arrays, i.e. crd / vals slices that start at arbitrary, unaligned elements) goes through the C ABI and the
    concatenated results must equal the oracle on the whole operand -- bit for bit, sharding must not change any
    row's operation order.  SpMV, SpMM, SDDMM, MTTKRP, SpAdd (C stays sharded) and SpGEMM (A sharded, B replicated).
  * with two or more GPUs, world-size-2 NCCL processes (one per GPU) do the same and all-gather the dense result rows
    (`partition.allgather_rows`), including two steps of the iterative SpMV  x <- A x  the all-gather exists for.
"""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "oracle"), HERE]

pytestmark = pytest.mark.gpu


def _dev(w):
    import gpu_util as G
    return G.to_device(w)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_csr_shards_through_the_cuda_kernels(world):
    import torch
    import oracle
    import synth
    import gpu_util as G
    from taco_b200 import partition
    w = synth.make("spmm", None, scale=13, K=32, dtype="float64")        # power-law rows: shards differ in row count
    n, m, K = w["dims"]
    x = synth.dense(synth.backend(None), m, 1, 77, np.float64)
    wd = _dev(w)
    xd = torch.as_tensor(x).cuda()
    bounds = partition.row_bounds(wd["A_pos"], n, world)
    assert bounds[0] == 0 and bounds[-1] == n
    ys, cs = [], []
    for r in range(world):
        sh = partition.shard_csr(wd["A_pos"], wd["A_crd"], wd["A_vals"], n, r, world, bounds)     # crd / vals are VIEWS
        rows = sh["row_end"] - sh["row_begin"]
        ys.append(G.run("spmv", dict(dims=[rows, m], A_pos=sh["pos"], A_crd=sh["crd"], A_vals=sh["vals"], x=xd)))
        cs.append(G.run("spmm", dict(dims=[rows, m, K], A_pos=sh["pos"], A_crd=sh["crd"], A_vals=sh["vals"], B=wd["B"])))
    y = np.concatenate(ys)
    C = np.concatenate(cs).reshape(n, K)
    want_y = oracle.spmv(w["A_pos"], w["A_crd"], w["A_vals"], x)
    # a row's nonzeros are summed in position order whatever shard it lands in; only the rows the kernels split
    # (SpMV: past two 2048-nonzero tiles, SpMM: long rows, > 128 nonzeros) are reassociated
    deg = np.diff(w["A_pos"])
    assert np.array_equal(y[deg <= 2048], want_y[deg <= 2048])
    assert np.allclose(y, want_y, rtol=1e-12, atol=0)
    want_C = oracle.spmm(w["A_pos"], w["A_crd"], w["A_vals"], w["B"].reshape(m, K))
    assert np.array_equal(C[deg <= 128], want_C[deg <= 128])
    assert np.allclose(C, want_C, rtol=1e-12, atol=0)


@pytest.mark.parametrize("world", [2, 5])
def test_sparse_output_shards(world):
    import oracle
    import synth
    import gpu_util as G
    from taco_b200 import partition
    # SpAdd: A, B and C share the row ranges, C stays sharded (pos rebased per shard).  SpGEMM: A sharded, B replicated.
    w = synth.make("spadd", None, n=30_011, deg=9, dtype="float64")
    n = w["dims"][0]
    wd = _dev(w)
    bounds = partition.row_bounds(wd["A_pos"], n, world)
    cp, cc, cv = oracle.spadd(w["A_pos"], w["A_crd"], w["A_vals"], w["B_pos"], w["B_crd"], w["B_vals"])
    gp, gc, gv = oracle.spgemm(w["A_pos"], w["A_crd"], w["A_vals"], w["B_pos"], w["B_crd"], w["B_vals"], n)
    add_parts, mul_parts = [], []
    for r in range(world):
        a = partition.shard_csr(wd["A_pos"], wd["A_crd"], wd["A_vals"], n, r, world, bounds)
        b = partition.shard_csr(wd["B_pos"], wd["B_crd"], wd["B_vals"], n, r, world, bounds)
        rows = a["row_end"] - a["row_begin"]
        add_parts.append(G.run("spadd", dict(dims=[rows, n], A_pos=a["pos"], A_crd=a["crd"], A_vals=a["vals"],
                                             B_pos=b["pos"], B_crd=b["crd"], B_vals=b["vals"])))
        mul_parts.append(G.run("spgemm", dict(dims=[rows, n, n], A_pos=a["pos"], A_crd=a["crd"], A_vals=a["vals"],
                                              B_pos=wd["B_pos"], B_crd=wd["B_crd"], B_vals=wd["B_vals"])))
    for parts, (wp, wc, wv) in ((add_parts, (cp, cc, cv)), (mul_parts, (gp, gc, gv))):
        pos = [np.zeros(1, np.int64)]
        for p, _, _ in parts:
            pos.append(p[1:].astype(np.int64) + pos[-1][-1])          # shard-local pos -> global
        assert np.array_equal(np.concatenate(pos), wp), "pos must be bit-exact"
        assert np.array_equal(np.concatenate([c for _, c, _ in parts]), wc), "crd must be bit-exact"
        assert np.array_equal(np.concatenate([v for _, _, v in parts]), wv)


@pytest.mark.parametrize("world", [2, 7])
def test_sddmm_and_csf_shards(world):
    import oracle
    import synth
    import gpu_util as G
    from taco_b200 import partition
    w = synth.make("sddmm", None, n=20_003, deg=12, K=32, dtype="float32")
    n, _, K = w["dims"]
    wd = _dev(w)
    bounds = partition.row_bounds(wd["B_pos"], n, world)
    _, _, want = oracle.sddmm(w["B_pos"], w["B_crd"], w["B_vals"], w["C"].reshape(n, K), w["D"].reshape(n, K))
    vals = []
    for r in range(world):
        b = partition.shard_csr(wd["B_pos"], wd["B_crd"], wd["B_vals"], n, r, world, bounds)
        r0, r1 = b["row_begin"], b["row_end"]
        _, _, v = G.run("sddmm", dict(dims=[r1 - r0, n, K], B_pos=b["pos"], B_crd=b["crd"], B_vals=b["vals"],
                                      C=wd["C"][r0 * K: r1 * K].contiguous(), D=wd["D"]))
        vals.append(v)
    assert np.allclose(np.concatenate(vals), want, rtol=1e-5, atol=0)
    # CSF mode-0 slice shards: each one is a self-contained MTTKRP over its own rows of A
    t = synth.make("mttkrp", None, I=4000, K=300, L=250, nnz=90_000, R=16, dtype="float64")
    I, Kd, L, R = t["dims"]
    td = _dev(t)
    want = oracle.mttkrp(t, t["C"].reshape(Kd, R), t["D"].reshape(L, R), I)
    blocks, nxt = [], 0
    for r in range(world):
        st = partition.shard_csf3(td, r, world, rebase_rows=True, dim0=I)
        assert st["row_begin"] == nxt
        nxt = st["row_end"]
        sub = {k: v for k, v in st.items() if k.startswith("B")}
        blocks.append(G.run("mttkrp", dict(dims=[st["row_end"] - st["row_begin"], Kd, L, R], C=td["C"], D=td["D"], **sub)))
    assert nxt == I
    assert np.array_equal(np.concatenate(blocks).reshape(I, R), want)


def test_result_fanout_into_mirror_windows_on_one_gpu():
    """taco_b200_set_result_peers on ONE GPU: two more windows of the same device stand in for the peers.  Every rank's shard
    is computed with its rows inside the gathered buffer; the kernels must leave the identical gathered result in the local
    window and in both mirrors (SpMM short + long rows, row-major; MTTKRP whole slices + hub slices)."""
    import torch
    import oracle
    import synth
    import gpu_util as G
    import taco_b200 as tb
    from taco_b200 import partition
    world = 3
    w = synth.make("spmm", None, scale=12, K=32, dtype="float64")
    n, m, K = w["dims"]
    wd = _dev(w)
    deg = np.diff(w["A_pos"])
    assert (deg > 128).any(), "the operand must exercise the long-row schedule too"
    want = oracle.spmm(w["A_pos"], w["A_crd"], w["A_vals"], w["B"].reshape(m, K))
    bufs = [torch.full((n * K,), -1.0, dtype=torch.float64, device="cuda") for _ in range(3)]
    bounds = partition.row_bounds(wd["A_pos"], n, world)
    tb.set_result_space("device")
    try:
        tb.set_result_peers(bufs[0].data_ptr(), [bufs[1].data_ptr(), bufs[2].data_ptr()], n * K * 8)
        for r in range(world):
            sh = partition.shard_csr(wd["A_pos"], wd["A_crd"], wd["A_vals"], n, r, world, bounds)
            rows = sh["row_end"] - sh["row_begin"]
            kk, tt = G.build("spmm", dict(dims=[rows, m, K], A_pos=sh["pos"], A_crd=sh["crd"], A_vals=sh["vals"], B=wd["B"]))
            tt[0].set_vals(bufs[0][sh["row_begin"] * K: sh["row_end"] * K])
            kk.compute(*tt)
        torch.cuda.synchronize()
        got = [b.cpu().numpy().reshape(n, K) for b in bufs]
        assert np.array_equal(got[0][deg <= 128], want[deg <= 128]) and np.allclose(got[0], want, rtol=1e-12, atol=0)
        assert np.array_equal(got[1], got[0]) and np.array_equal(got[2], got[0]), "a mirror window differs from the local result"
        # a result OUTSIDE the window must not fan out
        for b in bufs[1:]:
            b.fill_(-1.0)
        G.run("spmm", dict(dims=[n, m, K], A_pos=wd["A_pos"], A_crd=wd["A_crd"], A_vals=wd["A_vals"], B=wd["B"]))
        torch.cuda.synchronize()
        assert float(bufs[1].max()) == -1.0 and float(bufs[2].max()) == -1.0

        # fp32 (16-byte fragments of 4 columns) and a column-major result: the whole matrix inside the window
        w32 = synth.make("spmm", None, scale=11, K=64, dtype="float32")
        n2, m2, K2 = w32["dims"]
        wd32 = _dev(w32)
        want32 = oracle.spmm(w32["A_pos"], w32["A_crd"], w32["A_vals"], w32["B"].reshape(m2, K2))
        deg32 = np.diff(w32["A_pos"])
        for colmajor in (False, True):
            fb = [torch.full((n2 * K2,), -1.0, dtype=torch.float32, device="cuda") for _ in range(2)]
            tb.set_result_peers(fb[0].data_ptr(), [fb[1].data_ptr()], n2 * K2 * 4)
            kk, tt = G.build("spmm", dict(dims=[n2, m2, K2], A_pos=wd32["A_pos"], A_crd=wd32["A_crd"], A_vals=wd32["A_vals"], B=wd32["B"]),
                             colmajor_c=colmajor)
            tt[0].set_vals(fb[0])
            kk.compute(*tt)
            torch.cuda.synchronize()
            g0, g1 = (b.cpu().numpy().reshape((K2, n2) if colmajor else (n2, K2)) for b in fb)
            if colmajor:
                g0, g1 = g0.T, g1.T
            assert np.array_equal(g0[deg32 <= 128], want32[deg32 <= 128]) and np.allclose(g0, want32, rtol=1e-5, atol=1e-30)
            assert np.array_equal(g1, g0), "fp32 mirror window differs (colmajor=%s)" % colmajor

        # MTTKRP, including slices long enough for the slot-ordered hub chain (> 512 leaves)
        t = synth.make("mttkrp", None, I=64, K=300, L=250, nnz=60_000, R=16, dtype="float64")
        I, Kd, L, R = t["dims"]
        td = _dev(t)
        want_a = oracle.mttkrp(t, t["C"].reshape(Kd, R), t["D"].reshape(L, R), I)
        abufs = [torch.full((I * R,), -1.0, dtype=torch.float64, device="cuda") for _ in range(2)]
        tb.set_result_peers(abufs[0].data_ptr(), [abufs[1].data_ptr()], I * R * 8)
        for r in range(2):
            st = partition.shard_csf3(td, r, 2, rebase_rows=True, dim0=I)
            sub = {k: v for k, v in st.items() if k.startswith("B")}
            kk, tt = G.build("mttkrp", dict(dims=[st["row_end"] - st["row_begin"], Kd, L, R], C=td["C"], D=td["D"], **sub))
            tt[0].set_vals(abufs[0][st["row_begin"] * R: st["row_end"] * R])
            kk.compute(*tt)
        torch.cuda.synchronize()
        ga = [b.cpu().numpy().reshape(I, R) for b in abufs]
        assert np.allclose(ga[0], want_a, rtol=1e-12, atol=0)
        assert np.array_equal(ga[1], ga[0]), "the mirror window of the MTTKRP result differs from the local result"
    finally:
        tb.set_result_peers(None, None, 0)
        tb.set_result_space("host")


# ---------------------------------------------------------------------------------------------------------------
def _nccl_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TACO_B200_DEVICE=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle
        import synth
        import gpu_util as G
        import taco_b200 as tb
        from taco_b200 import partition
        tb.use_torch_stream()
        w = synth.make("spmm", None, scale=13, K=16, dtype="float64")
        n, m, K = w["dims"]
        wd = G.to_device(w)
        bounds = partition.row_bounds(wd["A_pos"], n, world)
        sh = partition.shard_csr(wd["A_pos"], wd["A_crd"], wd["A_vals"], n, rank, world, bounds)
        rows = sh["row_end"] - sh["row_begin"]
        # iterative SpMV  x <- A x : the all-gather of y between iterations is the path's only collective
        x = torch.as_tensor(synth.dense(synth.backend(None), m, 1, 5, np.float64)).cuda()
        x_ref = x.cpu().numpy()
        for _ in range(2):
            y_local = torch.as_tensor(G.run("spmv", dict(dims=[rows, m], A_pos=sh["pos"], A_crd=sh["crd"], A_vals=sh["vals"], x=x))).cuda()
            x = partition.allgather_rows(y_local, bounds)
            x_ref = oracle.spmv(w["A_pos"], w["A_crd"], w["A_vals"], x_ref)
        deg = np.diff(w["A_pos"])
        ok_spmv = bool(np.allclose(x.cpu().numpy(), x_ref, rtol=1e-11, atol=0))
        c_local = torch.as_tensor(G.run("spmm", dict(dims=[rows, m, K], A_pos=sh["pos"], A_crd=sh["crd"], A_vals=sh["vals"], B=wd["B"]))).cuda()
        C = partition.allgather_rows(c_local, bounds, row_len=K).cpu().numpy().reshape(n, K)
        want = oracle.spmm(w["A_pos"], w["A_crd"], w["A_vals"], w["B"].reshape(m, K))
        ok_spmm = bool(np.array_equal(C[deg <= 128], want[deg <= 128]) and np.allclose(C, want, rtol=1e-12, atol=0))
        t = synth.make("mttkrp", None, I=3000, K=200, L=150, nnz=40_000, R=8, dtype="float64")
        I, Kd, L, R = t["dims"]
        td = G.to_device(t)
        st = partition.shard_csf3(td, rank, world, rebase_rows=True, dim0=I)
        sub = {k: v for k, v in st.items() if k.startswith("B")}
        a_local = torch.as_tensor(G.run("mttkrp", dict(dims=[st["row_end"] - st["row_begin"], Kd, L, R], C=td["C"], D=td["D"], **sub))).cuda()
        ends = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
        dist.all_gather(ends, torch.tensor([st["row_end"]], dtype=torch.int64, device="cuda"))
        rb = np.array([0] + [int(e.item()) for e in ends])
        A = partition.allgather_rows(a_local, rb, row_len=R).cpu().numpy().reshape(I, R)
        ok_mttkrp = bool(np.array_equal(A, oracle.mttkrp(t, t["C"].reshape(Kd, R), t["D"].reshape(L, R), I)))
        # fused compute + all-gather: result rows stored to every GPU from inside the kernels, through peer-to-peer mappings
        # and through the NVLink multicast mapping of a symmetric allocation
        ok_fused = None
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(n * K, dtype=torch.float64, device="cuda")
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            abuf = symm.empty(I * R, dtype=torch.float64, device="cuda")
            ah = symm.rendezvous(abuf, dist.group.WORLD)
            have = True
        except Exception:  # noqa: BLE001
            have = False
        if have:
            want_a = oracle.mttkrp(t, t["C"].reshape(Kd, R), t["D"].reshape(L, R), I)

            def register(b_, h_, mode):
                if mode == "peers":
                    tb.set_result_peers(b_.data_ptr(), [int(p_) for r_, p_ in enumerate(h_.buffer_ptrs) if r_ != rank], b_.numel() * 8)
                else:
                    tb.set_result_multicast(b_.data_ptr(), int(h_.multicast_ptr), b_.numel() * 8)

            modes = ["peers"] + (["multicast"] if int(hdl.multicast_ptr) and int(ah.multicast_ptr) else [])
            ok_fused = True
            tb.set_result_space("device")
            try:
                for mode in modes:
                    buf.fill_(-1.0)
                    abuf.fill_(-1.0)
                    torch.cuda.synchronize()
                    hdl.barrier()
                    ah.barrier()
                    register(buf, hdl, mode)
                    kk, tt = G.build("spmm", dict(dims=[rows, m, K], A_pos=sh["pos"], A_crd=sh["crd"], A_vals=sh["vals"], B=wd["B"]))
                    tt[0].set_vals(buf[sh["row_begin"] * K: sh["row_end"] * K])
                    kk.compute(*tt)
                    hdl.barrier()
                    torch.cuda.synchronize()
                    got = buf.cpu().numpy().reshape(n, K)
                    ok_fused = ok_fused and bool(np.array_equal(got[deg <= 128], want[deg <= 128]) and np.allclose(got, want, rtol=1e-12, atol=0))
                    # MTTKRP rows through the same kind of window
                    register(abuf, ah, mode)
                    kk, tt = G.build("mttkrp", dict(dims=[st["row_end"] - st["row_begin"], Kd, L, R], C=td["C"], D=td["D"], **sub))
                    tt[0].set_vals(abuf[st["row_begin"] * R: st["row_end"] * R])
                    kk.compute(*tt)
                    ah.barrier()
                    torch.cuda.synchronize()
                    ok_fused = ok_fused and bool(np.array_equal(abuf.cpu().numpy().reshape(I, R), want_a))
            finally:
                tb.set_result_multicast(None, None, 0)
                tb.set_result_space("host")
        results[rank] = (ok_spmv, ok_spmm, ok_mttkrp, tb.launch_count() > 0, ok_fused)
    finally:
        dist.destroy_process_group()


def test_nccl_world2_sharded_kernels_and_allgather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2); the single-GPU shard tests above cover the kernels")
    import torch.multiprocessing as mp
    world = 2
    port = 29600 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_nccl_worker, args=(world, port, results), nprocs=world, join=True)
        assert len(results) == world
        for r in range(world):
            ok_spmv, ok_spmm, ok_mttkrp, launched, ok_fused = results[r]
            assert ok_fused is not False, "results stored to all GPUs from inside the kernels (peer stores / multicast) differ from the oracle on some rank"
            assert launched, "the CUDA kernels of libtaco_b200 must have run on every rank"
            assert ok_spmv, "iterative sharded SpMV + NCCL all-gather differs from the oracle"
            assert ok_spmm, "row-sharded SpMM + all-gather differs from the oracle"
            assert ok_mttkrp, "slice-sharded MTTKRP + all-gather differs from the oracle"
