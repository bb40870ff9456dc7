#!/usr/bin/env python
"""bench.py -- throughput of the taco GPU hot path on B200, next to the reference's CPU path on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload spmm|spmv|sddmm|mttkrp|spadd|spgemm|bspmm|bspmv|ttv|ttm|pack]
    python bench.py --impl reference ...          # the reference's own C/OpenMP codegen on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU; rank 0 prints the JSON line)

A "step" is one pass of the hot path over one batch of synthetic input.  Default workload = BASELINE.json configs[1]:
CSR SpMM C(i,k)=A(i,j)*B(j,k), fp32, power-law (R-MAT) 4Mi x 4Mi, 64Mi nnz, dense B with 128 columns.
  value      whole-job GFLOP/s with operands resident in HBM (device-resident taco_tensor_t, zero copies)
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of A and B, kernels, D2H of C per step
  roofline   dominant kernel: algorithmic bytes per launch / its CUDA-event duration, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  oracle/_ref (the real reference, JIT through cc, OpenMP on all host cores) on a bounded row slab
Multi-GPU: rows are independent, so each rank owns one row shard of the same size (weak scaling, B
replicated, no data-path collective); time = max over ranks of the device time.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FLOPS = {   # per step, as the reference counts them
    "spmv": lambda s: 2.0 * s["nnz"],
    "spmm": lambda s: 2.0 * s["nnz"] * s["K"],
    "sddmm": lambda s: 2.0 * s["nnz"] * s["K"],
    "mttkrp": lambda s: 3.0 * s["nnz"] * s["R"],
    "spadd": lambda s: 1.0 * s["nnzC"],
    "spgemm": lambda s: 2.0 * s["products"],
    "bspmm": lambda s: 2.0 * s["nnzb"] * s["br"] * s["bc"] * s["K"],
    "bspmv": lambda s: 2.0 * s["nnzb"] * s["br"] * s["bc"],
    "ttv": lambda s: 2.0 * s["nnz"],
    "ttm": lambda s: 2.0 * s["nnz"] * s["R"],
    "pack": lambda s: 1.0 * s["n"],            # not flops: coordinates packed (metric pack_gcoords, unit Gcoord/s)
}


def metric_of(wl):
    return (f"{wl}_gcoords", "Gcoord/s") if wl == "pack" else (f"{wl}_gflops", "GFLOP/s")
DOMINANT = {"spmv": "spmv_csr", "spmm": "spmm_csr", "sddmm": "sddmm_csr", "mttkrp": "mttkrp_csf",
            "spadd": "spadd_numeric", "spgemm": "spgemm_numeric", "bspmm": "bspmm_bcsr", "bspmv": "bspmv_bcsr", "ttv": "ttv_csf", "ttm": "ttm_csf", "pack": "pack_coo"}


def algorithmic_bytes(wl, s):
    """compulsory bytes of the dominant kernel per launch: every array touched once"""
    e = s["esize"]
    if wl == "spmv":
        return s["nnz"] * (4 + e) + 4 * (s["rows"] + 1) + e * (s["cols"] + s["rows"])
    if wl == "spmm":
        return s["nnz"] * (4 + e) + 4 * (s["rows"] + 1) + e * s["K"] * (s["cols"] + s["rows"])
    if wl == "sddmm":
        return s["nnz"] * (4 + 2 * e) + 4 * (s["rows"] + 1) + e * s["K"] * (s["cols"] + s["rows"])
    if wl == "mttkrp":
        return s["nnz"] * (4 + e) + 8 * s["nfib"] + 8 * s["nslices"] + e * s["R"] * (s["Kd"] + s["Ld"] + s["I"])
    if wl == "spadd":      # fused union kernel: both operands (crd + vals) read, result crd + vals written, all pos arrays
        return (s["nnzA"] + s["nnzB"]) * (4 + e) + s["nnzC"] * (4 + e) + 12 * (s["rows"] + 1)
    if wl == "spgemm":     # fill pass (sort + compress): A, the gathered B rows (crd + vals), the result, all pos arrays
        return (4 + e) * (s["nnzA"] + s["products"] + s["nnzC"]) + 12 * (s["rows"] + 1)
    if wl == "pack":       # coordinates and values read once, CSR arrays written once
        return s["n"] * (8 + e) + s["nnzC"] * (4 + e) + 4 * (s["rows"] + 1)
    if wl == "ttv":        # leaves + fiber / slice level arrays + c + the dense (I x K) result written once
        return s["nnz"] * (4 + e) + 8 * s["nfib"] + 8 * s["nslices"] + e * (s["Ld"] + s["I"] * s["Kd"])
    if wl == "ttm":        # + every row of C and of the (I*K x R) result touched once
        return s["nnz"] * (4 + e) + 8 * s["nfib"] + 8 * s["nslices"] + e * s["R"] * (s["Ld"] + s["I"] * s["Kd"])
    if wl == "bspmv":      # blocks streamed once, c and a touched once
        return 4 * (s["Mb"] + 1) + s["nnzb"] * (4 + e * s["br"] * s["bc"]) + e * (s["Nb"] * s["bc"] + s["Mb"] * s["br"])
    if wl == "bspmm":      # blocks streamed once, every row of B and C touched once
        return 4 * (s["Mb"] + 1) + s["nnzb"] * (4 + e * s["br"] * s["bc"]) + e * s["K"] * (s["Nb"] * s["bc"] + s["Mb"] * s["br"])
    raise KeyError(wl)


def sizes_of(wl, w, extra=None):
    d = [int(x) for x in w["dims"]]
    vals = w.get("A_vals", w.get("B_vals", w.get("vals")))
    e = 4 if "float32" in str(vals.dtype) else 8
    s = dict(esize=e)
    if wl in ("spmv", "spmm"):
        s.update(rows=d[0], cols=d[1], nnz=int(w["A_crd"].shape[0]), K=d[2] if wl == "spmm" else 1)
    elif wl == "sddmm":
        s.update(rows=d[0], cols=d[1], nnz=int(w["B_crd"].shape[0]), K=d[2])
    elif wl == "mttkrp":
        s.update(I=d[0], Kd=d[1], Ld=d[2], R=d[3], nnz=int(w["B3_crd"].shape[0]), nfib=int(w["B2_crd"].shape[0]),
                 nslices=int(w["B1_crd"].shape[0]))
    elif wl == "pack":
        s.update(rows=d[0], cols=d[1], n=int(w["vals"].shape[0]), nnzC=int(w["vals"].shape[0]))
    elif wl in ("ttv", "ttm"):
        s.update(I=d[0], Kd=d[1], Ld=d[2], R=d[3] if wl == "ttm" else 1, nnz=int(w["B3_crd"].shape[0]),
                 nfib=int(w["B2_crd"].shape[0]), nslices=int(w["B1_crd"].shape[0]))
    elif wl == "bspmv":
        s.update(Mb=d[0], Nb=d[1], br=d[2], bc=d[3], K=1, nnzb=int(w["A_crd"].shape[0]))
    elif wl == "bspmm":
        s.update(Mb=d[0], Nb=d[1], br=d[2], bc=d[3], K=d[4], nnzb=int(w["A_crd"].shape[0]))
    else:
        s.update(rows=d[0], nnzA=int(w["A_crd"].shape[0]), nnzB=int(w["B_crd"].shape[0]))
    if extra:
        s.update(extra)
    return s


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons with nvidia-smi during the timed region (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def ncu_traffic(wl, kernel_name):
    """dram__bytes_read+write per launch of the dominant kernel, from the committed ncu capture (profiles/traffic.json,
    written by tools/make_profiles.py from one `ncu --set full` run of this same command)"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(wl)
        return (t["dram_bytes_per_launch"], t["round"]) if t else (None, None)
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------
def make_workload(wl, device, rank, scale_down):
    from taco_b200 import synth
    over = {}
    if scale_down:    # quick functional runs (tests); never used for reported numbers
        over = {"spmm": dict(scale=16), "spmv": dict(n=100_000), "sddmm": dict(n=100_000),
                "mttkrp": dict(I=100_000, K=20_000, L=20_000, nnz=2_000_000), "spadd": dict(n=100_000),
                "spgemm": dict(n=50_000), "bspmm": dict(Mb=2048), "bspmv": dict(Mb=2048), "ttv": dict(I=512, K=512, L=50_000, nnz=2_000_000),
                "ttm": dict(I=128, K=128, L=50_000, nnz=1_000_000), "pack": dict(n=50_000, nnz=400_000)}[wl]
    if wl == "bspmm" and os.environ.get("TACO_B200_BENCH_BLOCK"):      # block-shape sweep (experiments only): 16 -> 16x16 blocks,
        b = int(os.environ["TACO_B200_BENCH_BLOCK"])                  # same matrix dimension and number of stored values
        over.update(br=b, bc=b, Mb=over.get("Mb", 32768) * 32 // b, deg=16 * 32 // b)
    old = synth.SEED0
    synth.SEED0 = old + 1000 * rank       # each rank owns a different row shard of the (N x larger) global operand
    try:
        w = synth.make(wl, device, **over)
        if rank and wl == "spmm":         # replicated dense operand: identical on every rank
            synth.SEED0 = old
            n = w["dims"][1]
            w["B"] = synth.dense(synth.backend(device), n, w["dims"][2], synth.SEED0 + 4, np.dtype("float32"))
        if rank and wl == "bspmm":
            synth.SEED0 = old
            w["B"] = synth.dense(synth.backend(device), w["dims"][1] * w["dims"][3], w["dims"][4], synth.SEED0 + 22,
                                 np.dtype("float32"))
    finally:
        synth.SEED0 = old
    return w


def reference_sample(wl, w, budget_rows):
    """bounded slab of the same workload for the CPU leg: the first `budget_rows` rows (slices) of the sparse operand,
    dense operands complete.  Returns (host arrays dict, fraction of the step's flops the slab represents)."""
    import gpu_util as G
    h = {}
    if wl in ("spmv", "spmm", "spadd", "spgemm"):
        rows = min(budget_rows, int(w["dims"][0]))
        pos = G.to_host(w["A_pos"][: rows + 1])
        nz = int(pos[-1])
        h.update(A_pos=pos, A_crd=G.to_host(w["A_crd"][:nz]), A_vals=G.to_host(w["A_vals"][:nz]))
        dims = [rows] + [int(x) for x in w["dims"][1:]]
        frac = nz / max(int(w["A_crd"].shape[0]), 1)
        if wl == "spmv":
            h["x"] = G.to_host(w["x"])
        elif wl == "spmm":
            h["B"] = G.to_host(w["B"])
        elif wl == "spadd":
            bpos = G.to_host(w["B_pos"][: rows + 1])
            bz = int(bpos[-1])
            h.update(B_pos=bpos, B_crd=G.to_host(w["B_crd"][:bz]), B_vals=G.to_host(w["B_vals"][:bz]))
        else:
            h.update(B_pos=G.to_host(w["B_pos"]), B_crd=G.to_host(w["B_crd"]), B_vals=G.to_host(w["B_vals"]))
    elif wl == "pack":       # the first budget_rows coordinates (same dimensions)
        n = min(budget_rows, int(w["vals"].shape[0]))
        h.update(c0=G.to_host(w["c0"][:n]), c1=G.to_host(w["c1"][:n]), vals=G.to_host(w["vals"][:n]))
        dims = [int(x) for x in w["dims"]]
        frac = n / max(int(w["vals"].shape[0]), 1)
    elif wl in ("bspmm", "bspmv"):
        rows = min(budget_rows, int(w["dims"][0]))
        pos = G.to_host(w["A_pos"][: rows + 1])
        nb = int(pos[-1])
        bsz = int(w["dims"][2]) * int(w["dims"][3])
        h.update(A_pos=pos, A_crd=G.to_host(w["A_crd"][:nb]), A_vals=G.to_host(w["A_vals"][: nb * bsz]))
        h.update(B=G.to_host(w["B"])) if wl == "bspmm" else h.update(c=G.to_host(w["c"]))
        dims = [rows] + [int(x) for x in w["dims"][1:]]
        frac = nb / max(int(w["A_crd"].shape[0]), 1)
    elif wl == "sddmm":
        rows = min(budget_rows, int(w["dims"][0]))
        pos = G.to_host(w["B_pos"][: rows + 1])
        nz = int(pos[-1])
        K = int(w["dims"][2])
        h.update(B_pos=pos, B_crd=G.to_host(w["B_crd"][:nz]), B_vals=G.to_host(w["B_vals"][:nz]),
                 C=G.to_host(w["C"][: rows * K]), D=G.to_host(w["D"]))
        dims = [rows, int(w["dims"][1]), K]
        frac = nz / max(int(w["B_crd"].shape[0]), 1)
    elif wl in ("mttkrp", "ttv", "ttm"):
        ns = min(budget_rows, int(w["B1_crd"].shape[0]))
        p2 = G.to_host(w["B2_pos"][: ns + 1])
        nf = int(p2[-1])
        p3 = G.to_host(w["B3_pos"][: nf + 1])
        nz = int(p3[-1])
        h.update(B1_pos=np.array([0, ns], np.int32), B1_crd=G.to_host(w["B1_crd"][:ns]), B2_pos=p2,
                 B2_crd=G.to_host(w["B2_crd"][:nf]), B3_pos=p3, B3_crd=G.to_host(w["B3_crd"][:nz]),
                 B_vals=G.to_host(w["B_vals"][:nz]))
        dims = [int(x) for x in w["dims"]]
        if wl == "mttkrp":
            h.update(C=G.to_host(w["C"]), D=G.to_host(w["D"]))
        else:               # dense (I x K [x R]) result: only the slab's leading rows
            dims[0] = int(h["B1_crd"][-1]) + 1 if ns else 1
            h.update(c=G.to_host(w["c"])) if wl == "ttv" else h.update(C=G.to_host(w["C"]))
        frac = nz / max(int(w["B3_crd"].shape[0]), 1)
    h["dims"] = np.array(dims, np.int32)
    return h, frac


def run_reference_cpu(wl, h, dtype, reps, threads):
    """times the reference's CPU implementation on a host slab: oracle/_ref harness (kind 'reference') when present,
    else the C oracle port.  Returns (best compute seconds incl. assemble for sparse outputs, kind)."""
    from taco_b200 import tbin
    harness = os.path.join(ROOT, "oracle", "_ref", "taco_ref_harness")
    sfx = "f32" if dtype == 4 else "f64"
    if os.path.exists(harness):
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") else None
        with tempfile.TemporaryDirectory(dir=tmp) as td:
            fin, fout = os.path.join(td, "in.tbin"), os.path.join(td, "out.tbin")
            tbin.write(fin, h)
            best = None
            for sched in (("cpu", "default") if wl != "sddmm" else ("default",)):
                r = subprocess.run([harness, "pack_csr" if wl == "pack" else wl, fin, fout, "--dtype", sfx, "--schedule", sched, "--threads", str(threads),
                                    "--reps", str(reps)], capture_output=True, text=True,
                                   env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
                if r.returncode != 0:
                    continue
                j = json.loads(r.stdout.strip().splitlines()[-1])
                t = [c + (a if wl in ("spadd", "spgemm") else 0.0) for a, c in zip(j["assemble_ms"], j["compute_ms"])]
                t = min(t[1:] if len(t) > 1 else t) / 1e3
                best = t if best is None else min(best, t)
            if best is not None:
                return best, "reference"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle.set_num_threads(threads)
    d = [int(x) for x in h["dims"]]
    fn = {
        "spmv": lambda: oracle.spmv(h["A_pos"], h["A_crd"], h["A_vals"], h["x"]),
        "spmm": lambda: oracle.spmm(h["A_pos"], h["A_crd"], h["A_vals"], h["B"].reshape(d[1], -1)),
        "sddmm": lambda: oracle.sddmm(h["B_pos"], h["B_crd"], h["B_vals"], h["C"].reshape(d[0], -1), h["D"].reshape(d[1], -1)),
        "mttkrp": lambda: oracle.mttkrp(h, h["C"].reshape(d[1], -1), h["D"].reshape(d[2], -1), d[0]),
        "spadd": lambda: oracle.spadd(h["A_pos"], h["A_crd"], h["A_vals"], h["B_pos"], h["B_crd"], h["B_vals"]),
        "spgemm": lambda: oracle.spgemm(h["A_pos"], h["A_crd"], h["A_vals"], h["B_pos"], h["B_crd"], h["B_vals"], d[-1]),
        "pack": lambda: oracle.pack("csr", d, [h["c0"], h["c1"]], h["vals"]),
        "ttv": lambda: oracle.ttv(h, h["c"], d[0], d[1]),
        "ttm": lambda: oracle.ttm(h, h["C"].reshape(d[2], -1), d[0], d[1]),
        "bspmv": lambda: oracle.bspmv(h["A_pos"], h["A_crd"], h["A_vals"].reshape(-1, d[2], d[3]), h["c"].reshape(d[1], d[3]), d[2], d[3]),
        "bspmm": lambda: oracle.bspmm(h["A_pos"], h["A_crd"], h["A_vals"].reshape(-1, d[2], d[3]),
                                      h["B"].reshape(d[1] * d[3], -1), d[2], d[3]),
    }[wl]
    best = None
    for _ in range(max(reps, 2)):
        t0 = time.perf_counter()
        fn()
        t = time.perf_counter() - t0
        best = t if best is None else min(best, t)
    return best, "port"


SAMPLE_ROWS = {"spmm": 1 << 19, "spmv": 1_000_000, "sddmm": 250_000, "mttkrp": 500_000, "spadd": 1_000_000,
               "spgemm": 200_000, "bspmm": 2048, "bspmv": 8192, "ttv": 1024, "ttm": 64, "pack": 2_000_000}


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="spmm", choices=sorted(FLOPS))
    ap.add_argument("--small", action="store_true", help="scaled-down operands (functional check only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = args.workload
    threads = os.cpu_count() or 1

    import torch
    import gpu_util as G

    if args.impl == "reference":
        if rank != 0:
            return 0
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        w = make_workload(wl, dev, 0, args.small)
        stats = sizes_of(wl, w)
        h, frac = reference_sample(wl, w, SAMPLE_ROWS[wl] if not args.small else 1 << 30)
        del w
        times = []
        kind = "port"
        for it in range(args.warmup + args.steps):
            t, kind = run_reference_cpu(wl, h, stats["esize"], 2, threads)
            if it >= args.warmup:
                times.append(t)
        sec = sum(times) / len(times)
        extra = {}
        if wl == "spadd":
            extra["nnzC"] = stats["nnzA"] + stats["nnzB"]
        if wl == "spgemm":
            extra["products"] = stats["nnzA"] * (stats["nnzB"] / max(stats["rows"], 1))
        stats.update(extra)
        gflops = FLOPS[wl](stats) * frac / sec / 1e9
        line = {"impl": "reference", "metric": metric_of(wl)[0], "value": gflops, "unit": metric_of(wl)[1], "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3 / max(frac, 1e-12),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32" if stats["esize"] == 4 else "f64", "data": "synthetic",
                "config": {"workload": workload_name(wl, stats), "sample": f"first {int(h['dims'][0])} rows"},
                "cpu_baseline": {"value": gflops, "unit": metric_of(wl)[1], "cores": threads, "kind": kind,
                                 "sample": f"first {int(h['dims'][0])} rows ({frac * 100:.1f}% of the nonzeros), "
                                           "time scaled to the full step"},
                "e2e": {"value": gflops, "unit": metric_of(wl)[1], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---- our arm ------------------------------------------------------------------------------------------------
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("TACO_B200_DEVICE", str(local_rank))
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import taco_b200 as tb
    from taco_b200 import _lib
    tb.use_torch_stream()
    tb.set_result_space("device")

    w = make_workload(wl, "cuda", rank, args.small)
    k, ts = G.build(wl, w)
    res = ts[0]
    stats = sizes_of(wl, w)
    sparse_out = wl in ("spadd", "spgemm", "sddmm", "pack")
    if not sparse_out:
        out = torch.empty(int(np.prod(res.dims)), dtype=torch.float32 if stats["esize"] == 4 else torch.float64, device="cuda")
        res.set_vals(out)

    def step():
        if sparse_out:
            k(*ts)          # GPU assembly (symbolic + scan + fill) and numeric phase: the whole sparse-output path
        else:
            k.compute(*ts)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if wl in ("spadd", "pack"):
        stats["nnzC"] = int(res.ct.vals_size)
    if wl == "spgemm":
        stats["nnzC"] = int(res.ct.vals_size)
        lens = (w["B_pos"][1:] - w["B_pos"][:-1]).to(torch.int64)
        stats["products"] = int(lens[w["A_crd"].to(torch.int64)].sum().item())
    small_inputs = algorithmic_bytes(wl, stats) < 4 * 126e6
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small_inputs else None

    launches0 = tb.launch_count()
    _lib.lib.taco_b200_profile_reset()
    _lib.lib.taco_b200_profile_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        if flush is not None:
            flush.fill_(1)          # evict L2 between timed iterations (operands smaller than ~4x L2)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    _lib.lib.taco_b200_profile_enable(0)
    launches = tb.launch_count() - launches0
    # the timed region of a ms-scale step is shorter than one nvidia-smi query: keep the identical loop running (untimed)
    # until the sampler has seen the load a few times, so the reported clocks / throttle reasons are those under this load
    t_more = time.perf_counter()
    while len(sampler.samples) < 6 and time.perf_counter() - t_more < 2.5:
        for _ in range(8):
            step()
        torch.cuda.synchronize()
    clocks = sampler.summary()
    clocks["sampling"] = "during the timed steps and an untimed continuation of the same loop (<= 2.5 s)"
    import ctypes
    kms, kn = ctypes.c_double(0), ctypes.c_int(0)
    _lib.lib.taco_b200_profile_get(DOMINANT[wl].encode(), ctypes.byref(kms), ctypes.byref(kn))
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    flops_rank = FLOPS[wl](stats)
    value = flops_rank * world / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    kern_ms = kms.value / max(kn.value, 1)
    ach = algorithmic_bytes(wl, stats) / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region ----------------------------
    e2e = None
    if not args.no_e2e:
        tb.set_result_space("host")
        hw, h2d = {}, 0
        for key, v in w.items():
            if key == "dims":
                hw[key] = v
                continue
            a = tb.pinned_empty(tuple(v.shape), G.np_dtype(v) if v.dtype.is_floating_point else np.int32)
            torch.from_numpy(a).copy_(v)
            hw[key] = a
            h2d += a.nbytes
        torch.cuda.synchronize()
        hk, hts = G.build(wl, hw)
        if sparse_out:
            d2h = None
        else:
            hout = tb.pinned_empty((int(np.prod(hts[0].dims)),), np.float32 if stats["esize"] == 4 else np.float64)
            hts[0].set_vals(hout)
            d2h = hout.nbytes
        e2e_steps = max(3, min(args.steps, 10))

        def hstep():
            if sparse_out:
                hk(*hts)
            else:
                hk.compute(*hts)

        hstep()
        if sparse_out:
            nn = int(hts[0].ct.vals_size)
            d2h = 4 * (stats["rows"] + 1) + nn * (4 + stats["esize"])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            hstep()             # returns after the D2H of the result has completed (host-visible result => sync)
        torch.cuda.synchronize()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": flops_rank * world / float(t.item()) / 1e9, "unit": metric_of(wl)[1], "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": float(t.item()) * 1e3, "steps": e2e_steps,
               "path": "taco_b200_<family>_compute(taco_tensor_t*) with pinned host arrays"}
        tb.set_result_space("device")

    # ---- the only collective of the path: all-gather of the dense result rows between iterations ----
    exchange = None
    if world > 1 and not sparse_out:
        gathered = torch.empty(world * out.numel(), dtype=out.dtype, device="cuda")
        for _ in range(2):
            dist.all_gather_into_tensor(gathered, out)
        torch.cuda.synchronize()
        dist.barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(5):
            dist.all_gather_into_tensor(gathered, out)
        eb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ea.elapsed_time(eb) / 5], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = out.numel() * out.element_size()
        exchange = {"collective": "all_gather_into_tensor (NCCL) of every rank's dense result rows, not part of the timed step",
                    "bytes_per_rank": nbytes, "ms": float(t.item()),
                    "recv_GBps_per_rank": nbytes * (world - 1) / (float(t.item()) * 1e-3) / 1e9}
        del gathered

    cpu = None
    if rank == 0 and not args.no_cpu:
        h, frac = reference_sample(wl, w, SAMPLE_ROWS[wl] if not args.small else 1 << 30)
        sec, kind = run_reference_cpu(wl, h, stats["esize"], 3, threads)
        cpu = {"value": flops_rank * frac / sec / 1e9, "unit": metric_of(wl)[1], "cores": threads, "kind": kind,
               "sample": (f"first {int(h['vals'].shape[0])} coordinates" if wl == "pack" else f"first {int(h['dims'][0])} rows") +
                         f" ({frac * 100:.1f}% of the nonzeros), best of 3"}

    if rank == 0:
        line = {"metric": metric_of(wl)[0], "value": value, "unit": metric_of(wl)[1], "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if stats["esize"] == 4 else "f64", "data": "synthetic",
                "config": {"workload": workload_name(wl, stats), "per_gpu": True,
                           "l2": "flushed between iterations" if flush is not None else "operands larger than L2, no flush",
                           "sharding": "row shard per rank, dense operand replicated, no collective"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": DOMINANT[wl], "achieved": ach, "peak": peak, "unit": "GB/s",
                             "frac": (ach / peak) if ach else None, "traffic": ncu_traffic(wl, DOMINANT[wl])[0],
                             "traffic_source": f"ncu --set full, profiles/{ncu_traffic(wl, DOMINANT[wl])[1]}_{wl}.md (full-size config)",
                             "peak_source": peak_src,
                             "kernel_ms": kern_ms, "algorithmic_bytes": algorithmic_bytes(wl, stats)},
                "cpu_baseline": cpu}
        if exchange:
            line["exchange"] = exchange
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def workload_name(wl, s):
    if wl == "spmm":
        return f"CSR SpMM fp32 R-MAT power-law {s['rows']}x{s['cols']} nnz={s['nnz']} K={s['K']}"
    if wl == "spmv":
        return f"CSR SpMV fp64 uniform {s['rows']}x{s['cols']} nnz={s['nnz']}"
    if wl == "sddmm":
        return f"CSR SDDMM fp32 uniform {s['rows']}x{s['cols']} nnz={s['nnz']} K={s['K']}"
    if wl == "bspmv":
        return (f"BCSR SpMV fp64 {s['Mb'] * s['br']}x{s['Nb'] * s['bc']} in {s['br']}x{s['bc']} blocks, {s['nnzb']} stored blocks")
    if wl == "bspmm":
        return (f"BCSR SpMM fp32 {s['Mb'] * s['br']}x{s['Nb'] * s['bc']} in {s['br']}x{s['bc']} blocks, "
                f"{s['nnzb']} stored blocks, K={s['K']}")
    if wl == "pack":
        return f"pack() COO -> CSR fp64 {s['rows']}x{s['cols']}, {s['n']} unsorted coordinates"
    if wl == "ttv":
        return f"CSF TTV fp64 {s['I']}x{s['Kd']}x{s['Ld']} nnz={s['nnz']} (dense {s['I']}x{s['Kd']} result)"
    if wl == "ttm":
        return f"CSF TTM fp64 {s['I']}x{s['Kd']}x{s['Ld']} nnz={s['nnz']} R={s['R']} (dense result)"
    if wl == "mttkrp":
        return f"CSF MTTKRP fp64 {s['I']}x{s['Kd']}x{s['Ld']} nnz={s['nnz']} R={s['R']}"
    return f"CSR {wl} fp64 {s['rows']} rows nnzA={s['nnzA']} nnzB={s['nnzB']} (GPU assembly + numeric)"


if __name__ == "__main__":
    sys.exit(main())
