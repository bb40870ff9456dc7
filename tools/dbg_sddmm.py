import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np, torch
import gpu_util as G, taco_b200 as tb
from taco_b200 import synth
for n in (100_000, 1_000_000, 2_000_000):
    w = synth.make("sddmm", "cuda", n=n)
    tb.set_result_space("device")
    k, ts = G.build("sddmm", w)
    k.assemble(*ts)
    tb.synchronize(); torch.cuda.synchronize()
    pos, crd = ts[0].level(1)
    print(n, "after assemble pos", pos[:6].tolist(), "ok", bool(torch.equal(pos, w["B_pos"])), "crd ok", bool(torch.equal(crd, w["B_crd"])), hex(pos.data_ptr()), hex(crd.data_ptr()), hex(ts[0].vals().data_ptr()))
    k.compute(*ts)
    tb.synchronize(); torch.cuda.synchronize()
    print(n, "after compute  pos", pos[:6].tolist(), "ok", bool(torch.equal(pos, w["B_pos"])), "crd ok", bool(torch.equal(crd, w["B_crd"])))
    k(*ts)
    tb.synchronize(); torch.cuda.synchronize()
    pos, crd = ts[0].level(1)
    print(n, "after evaluate pos", pos[:6].tolist(), "ok", bool(torch.equal(pos, w["B_pos"])), "crd ok", bool(torch.equal(crd, w["B_crd"])))
    tb.set_result_space("host")
