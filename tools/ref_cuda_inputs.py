"""Write the sparse operand of the C1 (SpMV) and C2 (SpMM) bench workloads as tbin files for tools/ref_cuda_bench.cu."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth
import tbin
out = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
for wl in (() if len(sys.argv) > 2 else ("spmv", "spmm")):
    w = synth.make(wl, "cuda")
    h = {k: (np.array(w[k][:2], np.int32) if k == "dims" else w[k].cpu().numpy()) for k in ("dims", "A_pos", "A_crd", "A_vals")}
    tbin.write(os.path.join(out, f"ref_{wl}.tbin"), h)
    print(wl, h["dims"], h["A_crd"].shape, h["A_vals"].dtype)
for wl in [a for a in sys.argv[2:] if a in ("mttkrp", "ttv", "ttm")]:
    w = synth.make(wl, "cuda")
    keys = ["B1_pos", "B1_crd", "B2_pos", "B2_crd", "B3_pos", "B3_crd", "B_vals"]
    h = {k: w[k].cpu().numpy() for k in keys}
    h["dims"] = np.array(w["dims"], np.int32)
    tbin.write(os.path.join(out, f"ref_{wl}.tbin"), h)
    print(wl, h["dims"], h["B3_crd"].shape)
