#!/usr/bin/env python
"""bench.py -- throughput of the taco GPU hot path on B200, next to the reference's CPU path on the same box.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # headline SpMM (C2) + SpMV (C1) + MTTKRP (C4)
    python bench.py --workload spmm|spmv|sddmm|mttkrp|spadd|spgemm|bspmm|bspmv|ttv|ttm|pack|mttkrp_fibers ...
    python bench.py --impl reference ...          # the reference's own C/OpenMP codegen on the host cores, full config
    torchrun --nproc-per-node N ... bench.py --gpus N ...   (one rank per GPU; rank 0 prints the JSON line)

A "step" is one pass of the hot path over the synthetic operands of a BASELINE.json config.  The default run measures
the three workloads the headline metric names; the JSON line's top-level keys are configs[1] (CSR SpMM fp32 power-law
4Mi x 4Mi, 64Mi nnz, K = 128), `workloads` holds the same record for SpMV (configs[0]) and MTTKRP (configs[3]).
  value      whole-job GFLOP/s with operands resident in HBM (device-resident taco_tensor_t, zero copies)
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of the operands, kernels, D2H of the result
  roofline   dominant kernel(s): algorithmic bytes per launch / CUDA-event duration, vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  oracle/_ref (the real reference, JIT through cc, OpenMP on all host cores) on a bounded row slab
Multi-GPU is STRONG scaling of the named shape: rank 0 generates the operand once, broadcasts it over
NCCL, every rank keeps its nnz-balanced shard (taco_b200.partition: CSR row ranges / CSF mode-0 slice ranges); dense
operands are replicated.  `value` = flops of the WHOLE config / max-over-ranks device time of the sharded kernels (the
path itself has no collective).  `iteration` = the same with the all-gather of the dense result rows (the exchange an
iterative caller needs: y -> x, C -> B, the updated factor) inside the step, overlapped with compute by row chunks.
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

FLOPS = {   # per step, as the reference counts them
    "spmv": lambda s: 2.0 * s["nnz"],
    "spmm": lambda s: 2.0 * s["nnz"] * s["K"],
    "sddmm": lambda s: 2.0 * s["nnz"] * s["K"],
    "mttkrp": lambda s: 3.0 * s["nnz"] * s["R"],
    "mttkrp_fibers": lambda s: 3.0 * s["nnz"] * s["R"],
    "spadd": lambda s: 1.0 * s["nnzC"],
    "spgemm": lambda s: 2.0 * s["products"],
    "bspmm": lambda s: 2.0 * s["nnzb"] * s["br"] * s["bc"] * s["K"],
    "bspmv": lambda s: 2.0 * s["nnzb"] * s["br"] * s["bc"],
    "ttv": lambda s: 2.0 * s["nnz"],
    "ttm": lambda s: 2.0 * s["nnz"] * s["R"],
    "pack": lambda s: 1.0 * s["n"],            # not flops: coordinates packed (metric pack_gcoords, unit Gcoord/s)
}
HEADLINE = ["spmm", "spmv", "mttkrp"]          # the workloads BASELINE.json's metric names
SHARDED = ("spmv", "spmm", "sddmm", "mttkrp", "mttkrp_fibers", "spadd", "spgemm")   # SURVEY.md 8(e); the rest run as replicas


def family_of(wl):
    return "mttkrp" if wl == "mttkrp_fibers" else wl


def metric_of(wl):
    return (f"{wl}_gcoords", "Gcoord/s") if wl == "pack" else (f"{wl}_gflops", "GFLOP/s")


DOMINANT = {"spmv": "spmv_csr", "spmm": "spmm_csr", "sddmm": "sddmm_csr", "mttkrp": "mttkrp_csf", "mttkrp_fibers": "mttkrp_csf",
            "spadd": "spadd_numeric", "spgemm": "spgemm_numeric", "bspmm": "bspmm_bcsr", "bspmv": "bspmv_bcsr",
            "ttv": "ttv_csf", "ttm": "ttm_csf", "pack": "pack_coo"}


def algorithmic_bytes(wl, s):
    """compulsory bytes of the dominant kernel per launch: every array touched once"""
    e = s["esize"]
    if wl == "spmv":
        return s["nnz"] * (4 + e) + 4 * (s["rows"] + 1) + e * (s["cols"] + s["rows"])
    if wl == "spmm":
        return s["nnz"] * (4 + e) + 4 * (s["rows"] + 1) + e * s["K"] * (s["cols"] + s["rows"])
    if wl == "sddmm":
        return s["nnz"] * (4 + 2 * e) + 4 * (s["rows"] + 1) + e * s["K"] * (s["cols"] + s["rows"])
    if wl in ("mttkrp", "mttkrp_fibers"):
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
    elif wl in ("mttkrp", "mttkrp_fibers"):
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
    if wl == "mttkrp_fibers":
        return (f"CSF MTTKRP fp64 {s['I']}x{s['Kd']}x{s['Ld']} nnz={s['nnz']} R={s['R']}, {s['nfib']} fibers "
                f"({s['nnz'] / max(s['nfib'], 1):.1f} leaves per fiber)")
    return f"CSR {wl} fp64 {s['rows']} rows nnzA={s['nnzA']} nnzB={s['nnzB']} (GPU assembly + numeric)"


def input_bytes(wl, s):
    """bytes of the operands a step reads (decides whether L2 is flushed between timed iterations)"""
    e = s["esize"]
    if wl in ("spadd", "spgemm"):
        return (s["nnzA"] + s["nnzB"]) * (4 + e) + 8 * (s["rows"] + 1)
    if wl == "pack":
        return s["n"] * (8 + e)
    return algorithmic_bytes(wl, s)


def config_of(wl, stats, world):
    """identical for both arms (the driver compares it): the named workload and how the timed region treats caches / ranks"""
    small = input_bytes(wl, stats) < 4 * 126e6
    return {"workload": workload_name(wl, stats),
            "l2": "flushed between iterations (256 MB write)" if small else "operands larger than L2, no flush",
            "sharding": ("single GPU" if world == 1 else
                         (f"strong scaling: the named operand nnz-split into {world} contiguous row / slice shards, dense operands "
                          "replicated, no collective inside the step" if wl in SHARDED else f"{world} independent replicas"))}


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """samples SM clocks / throttle reasons with nvidia-smi during the timed regions (B200_PROFILING.md recipe)"""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.active = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            if self.active.is_set():
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
                "reasons": reasons, "samples": len(self.samples),
                "sampling": "during the timed steps of every workload and an untimed continuation of the same loops (<= 2.5 s each)"}


def ncu_traffic(wl, prof_name):
    """dram__bytes_read+write per step of the dominant kernel(s), from the committed ncu capture (profiles/traffic.json,
    written by tools/make_profiles.py from one `ncu` run of this same command).  Refused (None) when the capture is of a
    different kernel set than the one this run launched."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(wl)
        if not t:
            return None, None
        if t.get("prof_name", prof_name) != prof_name:
            return None, f"profiles/traffic.json holds {t.get('prof_name')}, this run launched {prof_name}: stale, not reported"
        return t["dram_bytes_per_launch"], (f"ncu, profiles/{t['round']}_{wl}.md (full-size config, N=1, kernels {t.get('kernel')}; "
                                            f"captured at commit {t.get('commit', 'n/a')})")
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------------
SMALL = {"spmm": dict(scale=16), "spmv": dict(n=100_000), "sddmm": dict(n=100_000),
         "mttkrp": dict(I=100_000, K=20_000, L=20_000, nnz=2_000_000), "mttkrp_fibers": dict(I=20_000, K=20_000, L=20_000, nnz=2_000_000, nfib=160_000),
         "spadd": dict(n=100_000), "spgemm": dict(n=50_000), "bspmm": dict(Mb=2048), "bspmv": dict(Mb=2048),
         "ttv": dict(I=512, K=512, L=50_000, nnz=2_000_000), "ttm": dict(I=128, K=128, L=50_000, nnz=1_000_000),
         "pack": dict(n=50_000, nnz=400_000)}


def make_workload(wl, device, scale_down):
    """the named operand (tests/synth.py: counter-based hash, identical on every device)"""
    import synth
    over = dict(SMALL[wl]) if scale_down else {}   # scaled-down operands: functional checks only, never reported numbers
    if wl == "bspmm" and os.environ.get("TACO_B200_BENCH_BLOCK"):      # block-shape sweep (experiments only)
        b = int(os.environ["TACO_B200_BENCH_BLOCK"])
        over.update(br=b, bc=b, Mb=over.get("Mb", 32768) * 32 // b, deg=16 * 32 // b)
    return synth.make(wl, device, **over)


def to_host(a):
    if hasattr(a, "detach"):
        return a.detach().cpu().numpy()
    return np.asarray(a)


def reference_sample(wl, w, budget_rows):
    """bounded slab of the same workload for the CPU leg: the first `budget_rows` rows (slices) of the sparse operand,
    dense operands complete (budget_rows = None: the whole operand).  Returns (host arrays dict, fraction of the step's
    flops the slab represents)."""
    h = {}
    fam = family_of(wl)
    if budget_rows is None:
        budget_rows = 1 << 62
    if fam in ("spmv", "spmm", "spadd", "spgemm"):
        rows = min(budget_rows, int(w["dims"][0]))
        pos = to_host(w["A_pos"][: rows + 1])
        nz = int(pos[-1])
        h.update(A_pos=pos, A_crd=to_host(w["A_crd"][:nz]), A_vals=to_host(w["A_vals"][:nz]))
        dims = [rows] + [int(x) for x in w["dims"][1:]]
        frac = nz / max(int(w["A_crd"].shape[0]), 1)
        if fam == "spmv":
            h["x"] = to_host(w["x"])
        elif fam == "spmm":
            h["B"] = to_host(w["B"])
        elif fam == "spadd":
            bpos = to_host(w["B_pos"][: rows + 1])
            bz = int(bpos[-1])
            h.update(B_pos=bpos, B_crd=to_host(w["B_crd"][:bz]), B_vals=to_host(w["B_vals"][:bz]))
        else:
            h.update(B_pos=to_host(w["B_pos"]), B_crd=to_host(w["B_crd"]), B_vals=to_host(w["B_vals"]))
    elif fam == "pack":       # the first budget_rows coordinates (same dimensions)
        n = min(budget_rows, int(w["vals"].shape[0]))
        h.update(c0=to_host(w["c0"][:n]), c1=to_host(w["c1"][:n]), vals=to_host(w["vals"][:n]))
        dims = [int(x) for x in w["dims"]]
        frac = n / max(int(w["vals"].shape[0]), 1)
    elif fam in ("bspmm", "bspmv"):
        rows = min(budget_rows, int(w["dims"][0]))
        pos = to_host(w["A_pos"][: rows + 1])
        nb = int(pos[-1])
        bsz = int(w["dims"][2]) * int(w["dims"][3])
        h.update(A_pos=pos, A_crd=to_host(w["A_crd"][:nb]), A_vals=to_host(w["A_vals"][: nb * bsz]))
        h.update(B=to_host(w["B"])) if fam == "bspmm" else h.update(c=to_host(w["c"]))
        dims = [rows] + [int(x) for x in w["dims"][1:]]
        frac = nb / max(int(w["A_crd"].shape[0]), 1)
    elif fam == "sddmm":
        rows = min(budget_rows, int(w["dims"][0]))
        pos = to_host(w["B_pos"][: rows + 1])
        nz = int(pos[-1])
        K = int(w["dims"][2])
        h.update(B_pos=pos, B_crd=to_host(w["B_crd"][:nz]), B_vals=to_host(w["B_vals"][:nz]),
                 C=to_host(w["C"][: rows * K]), D=to_host(w["D"]))
        dims = [rows, int(w["dims"][1]), K]
        frac = nz / max(int(w["B_crd"].shape[0]), 1)
    elif fam in ("mttkrp", "ttv", "ttm"):
        ns = min(budget_rows, int(w["B1_crd"].shape[0]))
        p2 = to_host(w["B2_pos"][: ns + 1])
        nf = int(p2[-1])
        p3 = to_host(w["B3_pos"][: nf + 1])
        nz = int(p3[-1])
        h.update(B1_pos=np.array([0, ns], np.int32), B1_crd=to_host(w["B1_crd"][:ns]), B2_pos=p2,
                 B2_crd=to_host(w["B2_crd"][:nf]), B3_pos=p3, B3_crd=to_host(w["B3_crd"][:nz]),
                 B_vals=to_host(w["B_vals"][:nz]))
        dims = [int(x) for x in w["dims"]]
        if fam == "mttkrp":
            h.update(C=to_host(w["C"]), D=to_host(w["D"]))
        else:               # dense (I x K [x R]) result: only the slab's leading rows
            dims[0] = int(h["B1_crd"][-1]) + 1 if ns else 1
            h.update(c=to_host(w["c"])) if fam == "ttv" else h.update(C=to_host(w["C"]))
        frac = nz / max(int(w["B3_crd"].shape[0]), 1)
    h["dims"] = np.array(dims, np.int32)
    return h, frac


def run_reference_cpu(wl, h, dtype, reps, threads):
    """times the reference's CPU implementation on host arrays: oracle/_ref harness (kind 'reference': the real reference,
    JIT through cc, OpenMP) when present, else the C oracle port.  Returns (per-rep seconds -- the first rep is dropped when
    reps > 1 --, kind).  Sparse-output kernels count assemble + compute."""
    import tbin
    fam = family_of(wl)
    harness = os.path.join(ROOT, "oracle", "_ref", "taco_ref_harness")
    sfx = "f32" if dtype == 4 else "f64"
    if os.path.exists(harness):
        need = sum(v.nbytes for v in h.values()) + (64 << 20)
        tmp = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need else None
        with tempfile.TemporaryDirectory(dir=tmp) as td:
            fin, fout = os.path.join(td, "in.tbin"), os.path.join(td, "out.tbin")
            tbin.write(fin, h)
            best = None
            for sched in (("cpu", "default") if fam != "sddmm" else ("default",)):
                cmd = [harness, "pack_csr" if fam == "pack" else fam, fin, fout, "--dtype", sfx, "--schedule", sched,
                       "--threads", str(threads), "--reps", str(reps), "--no-out", "1"]
                r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS=str(threads)))
                if r.returncode != 0:
                    continue
                j = json.loads(r.stdout.strip().splitlines()[-1])
                t = [(c + (a if fam in ("spadd", "spgemm") else 0.0)) / 1e3 for a, c in zip(j["assemble_ms"], j["compute_ms"])]
                t = t[1:] if len(t) > 1 else t
                if best is None or sum(t) / len(t) < sum(best) / len(best):
                    best = t
            if best is not None:
                return best, "reference"
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
    }[fam]
    ts = []
    for _ in range(max(reps, 2)):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts[1:], "port"


# bounded slabs for the cpu_baseline leg of OUR arm (about 10-30 s of CPU work in total)
SAMPLE_ROWS = {"spmm": 1 << 19, "spmv": 1_000_000, "sddmm": 250_000, "mttkrp": 500_000, "mttkrp_fibers": 100_000, "spadd": 1_000_000,
               "spgemm": 200_000, "bspmm": 2048, "bspmv": 8192, "ttv": 1024, "ttm": 64, "pack": 2_000_000}
# the reference arm runs the FULL configuration of the headline workloads (None); the others keep a slab
REFERENCE_ROWS = dict(SAMPLE_ROWS, spmm=None, spmv=None, mttkrp=None, sddmm=None, spadd=None)


def derived_stats(wl, stats, w=None, res=None):
    """output-dependent sizes (sparse outputs): measured on our arm, estimated on the reference arm"""
    fam = family_of(wl)
    if fam in ("spadd", "pack") and res is not None:
        stats["nnzC"] = int(res.ct.vals_size)
    if fam == "spgemm" and res is not None:
        import torch
        stats["nnzC"] = int(res.ct.vals_size)
        lens = (w["B_pos"][1:] - w["B_pos"][:-1]).to(torch.int64)
        stats["products"] = int(lens[w["A_crd"].to(torch.int64)].sum().item())
    if res is None:
        if fam == "spadd":
            stats.setdefault("nnzC", stats["nnzA"] + stats["nnzB"])
        if fam == "spgemm":
            stats.setdefault("products", stats["nnzA"] * (stats["nnzB"] / max(stats["rows"], 1)))
            stats.setdefault("nnzC", int(stats["products"]))
    return stats


# ---------------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path (oracle/_ref), no product code loaded
# ---------------------------------------------------------------------------------------------------------------
def reference_record(wl, args, threads, steps, warmup):
    import torch
    dev = "cuda" if torch.cuda.is_available() else None
    w = make_workload(wl, dev, args.small)
    stats = derived_stats(wl, sizes_of(wl, w))
    rows = None if args.small else REFERENCE_ROWS[wl]
    if wl == "mttkrp" and rows is None:         # 5.1 GB of level arrays go through /dev/shm: keep a slab if it does not fit
        free = shutil.disk_usage("/dev/shm").free if os.path.isdir("/dev/shm") else 0
        if free < (7 << 30):
            rows = SAMPLE_ROWS[wl]
    h, frac = reference_sample(wl, w, rows)
    del w
    if dev:
        torch.cuda.empty_cache()
    ts, kind = run_reference_cpu(wl, h, stats["esize"], warmup + steps, threads)
    ts = ts[-steps:] if len(ts) >= steps else ts
    sec = sum(ts) / len(ts)
    value = FLOPS[wl](stats) * frac / sec / 1e9
    sample = ("the full configuration" if frac >= 0.999999 else
              f"first {int(h['dims'][0])} rows ({frac * 100:.1f}% of the nonzeros), time scaled to the full step")
    m, u = metric_of(wl)
    return {"metric": m, "value": value, "unit": u, "steps": len(ts), "warmup": warmup,
            "ms_per_step": sec * 1e3 / max(frac, 1e-12), "higher_is_better": True,
            "dtype": "f32" if stats["esize"] == 4 else "f64", "data": "synthetic",
            "config": config_of(wl, stats, max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": u, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": u, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
class Dist:
    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))

    def max(self, x):
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()


def broadcast_workload(w, D):
    """rank 0 generated the operand; ship every array to the other ranks over NCCL (NVLink) instead of regenerating it"""
    import torch
    import torch.distributed as dist
    if D.world == 1:
        return w
    keys = sorted(k for k in (w if D.rank == 0 else {}) if k != "dims")
    meta = [[int(x) for x in w["dims"]], [(k, tuple(w[k].shape), str(w[k].dtype)) for k in keys]] if D.rank == 0 else None
    box = [meta]
    dist.broadcast_object_list(box, src=0)
    dims, spec = box[0]
    out = {"dims": dims}
    for k, shape, dt in spec:
        t = w[k] if D.rank == 0 else torch.empty(shape, dtype=getattr(torch, dt.split(".")[-1]), device="cuda")
        dist.broadcast(t, src=0)
        out[k] = t
    return out


def shard_workload(wl, w, rank, world):
    """shard `rank` of `world` of an operand (taco_b200.partition: nnz-balanced contiguous row / slice ranges) as a
    self-contained workload dict, plus (rows of the dense result it owns, values per row, its first result row)"""
    import torch
    from taco_b200 import partition
    fam = family_of(wl)
    d = [int(x) for x in w["dims"]]
    rl = {"spmm": d[2] if fam == "spmm" else 1, "mttkrp": d[3] if fam == "mttkrp" else 1}.get(fam, 1)
    if world == 1 or wl not in SHARDED:
        return w, d[0], rl, 0
    own = lambda a: a.clone()       # the shard keeps its own arrays so the full operand can be released
    if fam in ("spmv", "spmm", "spgemm", "spadd", "sddmm"):
        p = "B" if fam == "sddmm" else "A"
        bounds = partition.row_bounds(w[f"{p}_pos"], d[0], world)
        sh = partition.shard_csr(w[f"{p}_pos"], w[f"{p}_crd"], w[f"{p}_vals"], d[0], rank, world, bounds)
        r0, r1 = sh["row_begin"], sh["row_end"]
        ws = {"dims": [r1 - r0] + d[1:], f"{p}_pos": own(sh["pos"]), f"{p}_crd": own(sh["crd"]), f"{p}_vals": own(sh["vals"])}
        if fam == "spmv":
            ws["x"] = w["x"]
        elif fam == "spmm":
            ws["B"] = w["B"]
        elif fam == "sddmm":
            ws["C"] = own(w["C"][r0 * d[2]: r1 * d[2]])
            ws["D"] = w["D"]
        elif fam == "spadd":
            sb = partition.shard_csr(w["B_pos"], w["B_crd"], w["B_vals"], d[0], rank, world, bounds)
            ws.update(B_pos=own(sb["pos"]), B_crd=own(sb["crd"]), B_vals=own(sb["vals"]))
        else:
            ws.update(B_pos=w["B_pos"], B_crd=w["B_crd"], B_vals=w["B_vals"])
        return ws, r1 - r0, rl, r0
    # CSF mode-0 slice shards: the shard owns the rows of A between its first slice and the next shard's first slice
    st = partition.shard_csf3(w, rank, world, rebase_rows=True, dim0=d[0])
    r0, r1 = st["row_begin"], st["row_end"]
    ws = {k: (own(v) if torch.is_tensor(v) else v) for k, v in st.items() if k.startswith("B")}
    ws.update(dims=[r1 - r0] + d[1:], C=w["C"], D=w["D"])
    return ws, r1 - r0, rl, r0


def host_copy(ws, tb, torch, G):
    """pinned host copies of a workload dict -> (host dict, bytes)"""
    hw, nbytes = {}, 0
    for key, v in ws.items():
        if key == "dims":
            hw[key] = v
            continue
        a = tb.pinned_empty(tuple(v.shape), G.np_dtype(v) if v.dtype.is_floating_point else np.int32)
        torch.from_numpy(a).copy_(v)
        hw[key] = a
        nbytes += a.nbytes
    return hw, nbytes


REPLICATED = {"spmv": ["x"], "spmm": ["B"], "sddmm": ["D"], "mttkrp": ["C", "D"]}    # dense operands every rank needs whole


def ours_record(wl, args, D, sampler):
    import torch
    import torch.distributed as dist
    import gpu_util as G
    import taco_b200 as tb
    from taco_b200 import _lib
    fam = family_of(wl)
    world = D.world
    sharded = world > 1 and wl in SHARDED
    threads = os.cpu_count() or 1
    tb.set_result_space("device")

    # ---- the named operand: generated once (rank 0), broadcast over NCCL, sharded -----------------------------------
    w = make_workload(wl, "cuda", args.small) if (D.rank == 0 or not sharded) else None
    if sharded:
        w = broadcast_workload(w, D)
    full = sizes_of(wl, w)
    ws, out_rows, row_len, row0 = shard_workload(wl, w, D.rank, world)
    total_rows = int(w["dims"][0])
    cpu_sample = None
    if D.rank == 0 and world == 1 and not args.no_cpu:
        cpu_sample = reference_sample(wl, w, SAMPLE_ROWS[wl] if not args.small else None)
    if sharded:
        del w
        torch.cuda.empty_cache()
    k, ts = G.build(fam, ws)
    res = ts[0]
    stats = sizes_of(wl, ws)
    sparse_out = fam in ("spadd", "spgemm", "sddmm", "pack")
    tdt = torch.float32 if stats["esize"] == 4 else torch.float64
    if not sparse_out:
        out = torch.empty(int(np.prod(res.dims)), dtype=tdt, device="cuda")
        res.set_vals(out)

    def step():
        if sparse_out:
            k(*ts)          # GPU assembly (symbolic + scan + fill) and numeric phase: the whole sparse-output path
        else:
            k.compute(*ts)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    derived_stats(wl, stats, ws, res)
    for key in ("nnzC", "products"):            # whole-job output-dependent sizes: sum over the shards
        if key in stats:
            t = torch.tensor([float(stats[key])], dtype=torch.float64, device="cuda")
            if sharded:
                dist.all_reduce(t)
            full[key] = float(t.item())
    cfg = config_of(wl, full, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if cfg["l2"].startswith("flushed") else None

    launches0 = tb.launch_count()
    _lib.lib.taco_b200_profile_reset()
    _lib.lib.taco_b200_profile_enable(1)
    sampler.active.set()
    D.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        if flush is not None:
            flush.fill_(1)          # evict L2 between timed iterations (operands smaller than ~4x L2)
        a.record()
        step()
        b.record()
    torch.cuda.synchronize()
    D.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    _lib.lib.taco_b200_profile_enable(0)
    launches = tb.launch_count() - launches0
    # the timed region of a ms-scale step is shorter than one nvidia-smi query: keep the identical loop running (untimed)
    # until the sampler has seen this load a few times, so the reported clocks / throttle reasons are those under load
    seen0, t_more = len(sampler.samples), time.perf_counter()
    while len(sampler.samples) < seen0 + 4 and time.perf_counter() - t_more < 2.5:
        for _ in range(8):
            step()
        torch.cuda.synchronize()
    sampler.active.clear()
    kms, kn = ctypes.c_double(0), ctypes.c_int(0)
    _lib.lib.taco_b200_profile_get(DOMINANT[wl].encode(), ctypes.byref(kms), ctypes.byref(kn))
    ms_per_step = D.max(dev_ms) / args.steps
    flops_job = FLOPS[wl](full) * (1 if (sharded or world == 1) else world)
    value = flops_job / (ms_per_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    kern_ms = kms.value / max(kn.value, 1)
    abytes = algorithmic_bytes(wl, stats)
    ach = abytes / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else None
    traffic, traffic_src = ncu_traffic(wl, DOMINANT[wl]) if world == 1 and not args.small else (None, "captured at N=1, full size only")

    # ---- iteration: the step plus the all-gather of the dense result rows, overlapped by row chunks ------------------
    iteration = None
    if sharded and not sparse_out:
        iteration = iteration_mode(wl, fam, ws, row_len, args, D, flops_job, tdt)
        iteration["fused"] = fused_iteration(wl, fam, ws, row_len, row0, out_rows, total_rows, args, D, flops_job, tdt)

    # ---- e2e: host (pinned) buffers through the C ABI, copies inside the timed region --------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = e2e_mode(wl, fam, ws, stats, args, D, flops_job, sparse_out, sharded)
    tb.set_result_space("device")

    cpu = None
    if cpu_sample is not None:
        h, frac = cpu_sample
        tsec, kind = run_reference_cpu(wl, h, stats["esize"], 3, threads)
        sec = min(tsec)
        what = (f"first {int(h['vals'].shape[0])} coordinates" if fam == "pack" else
                f"first {int(h['B1_crd'].shape[0])} slices" if fam in ("mttkrp", "ttv", "ttm") else f"first {int(h['dims'][0])} rows")
        cpu = {"value": FLOPS[wl](full) * frac / sec / 1e9, "unit": metric_of(wl)[1], "cores": threads, "kind": kind,
               "sample": what +
                         f" ({frac * 100:.1f}% of the nonzeros), best of {len(tsec)}"}
    m, u = metric_of(wl)
    rec = {"metric": m, "value": value, "unit": u, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if wl in SHARDED else "weak",
           "vs_baseline": None, "dtype": "f32" if stats["esize"] == 4 else "f64", "data": "synthetic", "config": cfg,
           "e2e": e2e, "gpu_launches": launches,
           "roofline": {"bound": "hbm", "kernel": DOMINANT[wl], "achieved": ach, "peak": peak, "unit": "GB/s",
                        "frac": (ach / peak) if ach else None, "traffic": traffic, "traffic_source": traffic_src,
                        "peak_source": peak_src, "kernel_ms": kern_ms, "algorithmic_bytes": abytes,
                        "scope": "rank 0's shard" if sharded else "the whole operand"},
           "cpu_baseline": cpu}
    if iteration:
        rec["iteration"] = iteration
    del k, ts, res, ws
    torch.cuda.empty_cache()
    return rec


def iteration_mode(wl, fam, ws, row_len, args, D, flops_job, tdt):
    """step = kernels on `nch` row chunks of the shard; chunk c's all-gather (NCCL over NVLink) runs on a side stream while
    chunk c+1 computes.  Uneven chunks are padded to the largest rank's chunk for the collective.  Results too small for
    chunking to pay (under 32 MB per rank: the collective is latency-bound) are gathered in one piece."""
    import torch
    import torch.distributed as dist
    import gpu_util as G
    world = D.world
    nch = 4 if int(ws["dims"][0]) * row_len * (4 if tdt == torch.float32 else 8) >= (32 << 20) else 1
    chunks = [shard_workload(wl, ws, c, nch)[:2] for c in range(nch)]
    rows_t = torch.tensor([r for _, r in chunks], dtype=torch.int64, device="cuda")
    allrows = [torch.zeros_like(rows_t) for _ in range(world)]
    dist.all_gather(allrows, rows_t)
    maxrows = torch.stack(allrows).max(dim=0).values.tolist()
    runs = []
    for c, (cw, rows_c) in enumerate(chunks):
        k, ts = G.build(fam, cw)
        send = torch.zeros(max(int(maxrows[c]), 1) * row_len, dtype=tdt, device="cuda")     # padded to the largest rank's chunk
        ts[0].set_vals(send[: rows_c * row_len])
        recv = torch.empty(world * send.numel(), dtype=tdt, device="cuda")
        runs.append((k, ts, send, recv, rows_c))
    comm = torch.cuda.Stream()
    main = torch.cuda.current_stream()

    def istep():
        works = []
        for k, ts, send, recv, rows_c in runs:
            if rows_c > 0:
                k.compute(*ts)
            e = torch.cuda.Event()
            e.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(e)
                works.append(dist.all_gather_into_tensor(recv, send, async_op=True))
        for wk in works:
            wk.wait()               # the compute stream waits for the collectives: the next iteration may read the gathered rows

    for _ in range(max(args.warmup, 3)):
        istep()
    torch.cuda.synchronize()
    D.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        istep()
    b.record()
    torch.cuda.synchronize()
    D.barrier()
    ms = D.max(a.elapsed_time(b)) / args.steps
    a.record()                      # the collective alone, for reference
    for _ in range(5):
        for k, ts, send, recv, rows_c in runs:
            dist.all_gather_into_tensor(recv, send)
    b.record()
    torch.cuda.synchronize()
    ag_ms = D.max(a.elapsed_time(b)) / 5
    nbytes = sum(send.numel() * send.element_size() for _, _, send, _, _ in runs)
    return {"value": flops_job / (ms * 1e-3) / 1e9, "unit": metric_of(wl)[1], "ms_per_step": ms,
            "collective": f"all_gather_into_tensor (NCCL) of every rank's dense result rows inside the step, {nch} row chunks, "
                          "chunk c gathered on a side stream while chunk c+1 computes",
            "allgather_alone_ms": ag_ms, "allgather_bytes_per_rank": nbytes,
            "allgather_recv_GBps_per_rank": nbytes * (world - 1) / (ag_ms * 1e-3) / 1e9}


def fused_iteration(wl, fam, ws, row_len, row0, out_rows, total_rows, args, D, flops_job, tdt):
    """the same iteration with the all-gather INSIDE the kernel: the dense result lives in a symmetric allocation
    (torch.distributed._symmetric_memory) and the library stores every result row to all GPUs while the kernel is still
    computing -- `peers`: a local store plus one peer-to-peer store per other GPU (taco_b200_set_result_peers); `multicast`: one
    store through the NVLink multicast mapping, replicated by the NVSwitch (taco_b200_set_result_multicast).  The step ends with a
    device-side cross-rank barrier.  No NCCL collective, no second pass over the result."""
    import torch
    import torch.distributed as dist
    import gpu_util as G
    import taco_b200 as tb
    if fam not in ("spmm", "mttkrp"):
        return {"unavailable": "in-kernel result fan-out is implemented for the SpMM and MTTKRP results"}
    try:
        import torch.distributed._symmetric_memory as symm
        buf = symm.empty(total_rows * row_len, dtype=tdt, device="cuda")
        hdl = symm.rendezvous(buf, dist.group.WORLD)
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"symmetric memory: {type(e).__name__}: {str(e)[:160]}"}
    nbytes = buf.numel() * buf.element_size()
    k, ts = G.build(fam, ws)
    ts[0].set_vals(buf[row0 * row_len: (row0 + out_rows) * row_len])

    def fstep():
        if out_rows > 0:
            k.compute(*ts)
        hdl.barrier()           # every rank's rows have landed everywhere

    def measure(what):
        buf.zero_()
        torch.cuda.synchronize()
        D.barrier()
        for _ in range(max(args.warmup, 3)):
            fstep()
        torch.cuda.synchronize()
        D.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fstep()
        b.record()
        torch.cuda.synchronize()
        D.barrier()
        ms = D.max(a.elapsed_time(b)) / args.steps
        # every rank must now hold the whole result: compare checksums of the gathered buffer across ranks
        chk = buf.double().sum().reshape(1)
        lo, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        return {"value": flops_job / (ms * 1e-3) / 1e9, "unit": metric_of(wl)[1], "ms_per_step": ms,
                "all_ranks_hold_identical_result": bool(lo.item() == hi_.item() and chk.item() != 0.0), "collective": what,
                "result_bytes_per_rank": out_rows * row_len * buf.element_size()}

    out = {}
    rank, world = D.rank, D.world
    try:
        peers = [int(p) for r, p in enumerate(hdl.buffer_ptrs) if r != rank]
        if 1 <= len(peers) <= 7:
            tb.set_result_peers(buf.data_ptr(), peers, nbytes)
            out["peers"] = measure(f"none: every result row stored locally and into the {world - 1} peer GPU(s) from inside the kernel "
                                   "(st.global over NVLink peer mappings), then a device-side cross-rank barrier")
            tb.set_result_peers(None, None, 0)
        mc = int(hdl.multicast_ptr)
        if mc:
            tb.set_result_multicast(buf.data_ptr(), mc, nbytes)
            out["multicast"] = measure("none: result rows stored through the NVLink multicast mapping inside the kernel (multimem.st), "
                                       "then a device-side cross-rank barrier")
        else:
            out["multicast"] = {"unavailable": "no NVLink multicast mapping for the symmetric allocation on this box"}
    finally:
        tb.set_result_multicast(None, None, 0)
    best = min((v for v in out.values() if "ms_per_step" in v), key=lambda v: v["ms_per_step"], default=None)
    if best is not None:
        out.update(value=best["value"], unit=best["unit"], ms_per_step=best["ms_per_step"])
    return out


def e2e_mode(wl, fam, ws, stats, args, D, flops_job, sparse_out, sharded):
    """the same step through the C ABI with HOST (pinned) buffers.  N = 1: every operand is a host array, the library
    stages it (H2D), runs the kernels and copies the result back (D2H).  N > 1: each rank uploads its shard of the sparse
    operand and 1/N of every replicated dense operand; the slices are all-gathered over NVLink (one PCIe upload of the dense
    operand per JOB instead of one per rank) and the library is called with host sparse arrays + the gathered device operand."""
    import torch
    import torch.distributed as dist
    import gpu_util as G
    import taco_b200 as tb
    world = D.world
    tb.set_result_space("host")
    repl = REPLICATED.get(fam, []) if sharded else []
    hw, h2d = host_copy({k2: v for k2, v in ws.items() if k2 not in repl}, tb, torch, G)
    gathered, pinned_slices = {}, []
    for name in repl:           # 1/N slice per rank in pinned host memory, all-gathered into a padded device buffer
        v = ws[name]
        per = (v.numel() + world - 1) // world
        lo, hi = min(D.rank * per, v.numel()), min((D.rank + 1) * per, v.numel())
        hs = tb.pinned_empty((per,), G.np_dtype(v))
        pinned_slices.append(hs)
        torch.from_numpy(hs)[: hi - lo].copy_(v[lo:hi])
        stage = torch.empty(per, dtype=v.dtype, device="cuda")
        fullbuf = torch.empty(per * world, dtype=v.dtype, device="cuda")
        gathered[name] = (torch.from_numpy(hs), stage, fullbuf)
        hw[name] = fullbuf[: v.numel()]
        h2d += hs.nbytes
    torch.cuda.synchronize()
    hk, hts = G.build(fam, hw)
    d2h = None
    if not sparse_out:
        hout = tb.pinned_empty((int(np.prod(hts[0].dims)),), np.float32 if stats["esize"] == 4 else np.float64)
        pinned_slices.append(hout)
        hts[0].set_vals(hout)
        d2h = hout.nbytes
    e2e_steps = max(3, min(args.steps, 10 if fam != "mttkrp" else 5))

    def hstep():
        for name, (hs, stage, fullbuf) in gathered.items():
            stage.copy_(hs, non_blocking=True)
            dist.all_gather_into_tensor(fullbuf, stage)
        if sparse_out:
            hk(*hts)
        else:
            hk.compute(*hts)

    hstep()
    if sparse_out:
        nn = int(hts[0].ct.vals_size)
        d2h = 4 * (int(hts[0].dims[0]) + 1) + nn * (4 + stats["esize"])
    D.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hstep()             # returns after the D2H of the result has completed (host-visible result => sync)
    torch.cuda.synchronize()
    e2e_s = D.max((time.perf_counter() - t0) / e2e_steps)
    pcie = pcie_probe(D, torch, tb) if world > 1 else None
    tot = torch.tensor([float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1 and sharded:
        dist.all_reduce(tot)
    elif world > 1:
        tot *= world
    del hk, hts
    for v in list(hw.values()) + pinned_slices:
        if isinstance(v, np.ndarray):
            tb.pinned_free(v)
    out = {} if pcie is None else {"pcie_probe": pcie}
    return {**out, "value": flops_job / e2e_s / 1e9, "unit": metric_of(wl)[1], "h2d_bytes_per_step": int(tot[0].item()),
            "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
            "path": "taco_b200_<family>_compute(taco_tensor_t*) with pinned host arrays" +
                    ("" if not repl else f"; replicated dense operand(s) {repl}: 1/{world} uploaded per rank + NCCL all-gather; bytes summed over ranks")}


def pcie_probe(D, torch, tb, mb=256):
    """what the box's host<->device path gives when every rank copies at once (256 MB up and 256 MB down per rank, pinned):
    the e2e number of a PCIe-bound step cannot scale past this aggregate, whatever the kernels do"""
    n = mb << 20
    h_up, h_dn = tb.pinned_empty((n,), np.uint8), tb.pinned_empty((n,), np.uint8)
    d_up, d_dn = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    t_up, t_dn = torch.from_numpy(h_up), torch.from_numpy(h_dn)
    side = torch.cuda.Stream()
    best = None
    for _ in range(3):
        D.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d_up.copy_(t_up, non_blocking=True)
        with torch.cuda.stream(side):
            t_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        t = D.max(time.perf_counter() - t0)
        best = t if best is None else min(best, t)
    tb.pinned_free(h_up)
    tb.pinned_free(h_dn)
    return {"aggregate_GBps": 2 * n * D.world / best / 1e9, "per_rank_GBps": 2 * n / best / 1e9,
            "what": f"{D.world} ranks copying {mb} MB up and {mb} MB down at once (pinned), max over ranks"}


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(FLOPS), help="one workload only (default: spmm + spmv + mttkrp)")
    ap.add_argument("--small", action="store_true", help="scaled-down operands (functional check only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    D = Dist()
    wls = [args.workload] if args.workload else list(HEADLINE)
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if D.rank != 0:
            return 0
        recs = []
        for wl in wls:
            heavy = wl in ("mttkrp", "mttkrp_fibers")
            recs.append(reference_record(wl, args, threads, max(min(args.steps, 5) if heavy else args.steps, 1),
                                         max(min(args.warmup, 1) if heavy else args.warmup, 0)))
        line = dict(recs[0])
        line.update({"impl": "reference", "n_gpus": args.gpus, "scaling": "strong", "vs_baseline": None,
                     "steps": args.steps, "warmup": args.warmup})
        if len(recs) > 1:
            line["workloads"] = {wl: r for wl, r in zip(wls[1:], recs[1:])}
        print(json.dumps(line))
        return 0

    import torch
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    args.warmup = max(args.warmup, 3)
    torch.cuda.set_device(D.local_rank)
    os.environ.setdefault("TACO_B200_DEVICE", str(D.local_rank))
    import torch.distributed as dist
    if D.world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", D.local_rank))
    import taco_b200 as tb
    tb.use_torch_stream()
    sampler = ClockSampler(D.local_rank)
    sampler.start()
    recs = [ours_record(wl, args, D, sampler) for wl in wls]
    clocks = sampler.summary()
    if D.rank == 0:
        line = dict(recs[0])
        line["clocks"] = clocks
        if len(recs) > 1:
            line["workloads"] = {wl: r for wl, r in zip(wls[1:], recs[1:])}
        print(json.dumps(line))
    if D.world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
