"""Where does the host-buffer (e2e) time of the sparse-output paths go?  python tools/e2e_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import torch
import gpu_util as G
import taco_b200 as tb
import synth

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

if os.environ.get("PROBE_TORCH_STREAM"):
    tb.use_torch_stream()
w = synth.make("pack", "cuda")
hw = {}
for k, v in w.items():
    if k == "dims": hw[k] = v; continue
    a = tb.pinned_empty(tuple(v.shape), np.float64 if v.dtype.is_floating_point else np.int32)
    torch.from_numpy(a).copy_(v); hw[k] = a
pw = {k: (v if k == "dims" else np.array(v)) for k, v in hw.items()}      # pageable copies
dims = list(w["dims"])
for space in ("device", "host"):
    tb.set_result_space(space)
    print(f"pack result={space:6s} inputs=device  : {t(lambda: tb.pack('A', dims, tb.CSR, [w['c0'], w['c1']], w['vals'])):8.2f} ms")
    print(f"pack result={space:6s} inputs=pinned  : {t(lambda: tb.pack('A', dims, tb.CSR, [hw['c0'], hw['c1']], hw['vals'])):8.2f} ms")
    print(f"pack result={space:6s} inputs=pageable: {t(lambda: tb.pack('A', dims, tb.CSR, [pw['c0'], pw['c1']], pw['vals'])):8.2f} ms")
# raw copies of the same volume for scale
d = torch.empty(124_000_000, dtype=torch.uint8, device="cuda")
hp = torch.empty(124_000_000, dtype=torch.uint8).pin_memory()
hq = torch.empty(124_000_000, dtype=torch.uint8)
print(f"D2H 124 MB pinned   : {t(lambda: hp.copy_(d)):8.2f} ms")
print(f"D2H 124 MB pageable : {t(lambda: hq.copy_(d)):8.2f} ms")
def fresh():
    x = np.empty(124_000_000, np.uint8); torch.from_numpy(x).copy_(d)
print(f"D2H 124 MB into a fresh malloc : {t(fresh):8.2f} ms")

import cProfile, pstats
tb.set_result_space("host")
pr = cProfile.Profile(); pr.enable()
for _ in range(3): tb.pack('A', dims, tb.CSR, [hw['c0'], hw['c1']], hw['vals'])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
