"""bench.py prints ONE JSON line that carries the driver's contract keys (`-m gpu`; scaled-down operands, so the numbers mean
nothing -- only the shape of the line and that every leg runs: device-resident step, e2e through host buffers, CPU leg)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
pytestmark = pytest.mark.gpu


def _line(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--small", "--steps", "3", "--warmup", "3"] + args,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("wl", ["spmm", "spadd", "mttkrp", "pack"])
def test_bench_line_contract(wl):
    j = _line(["--workload", wl])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["n_gpus"] == 1 and j["steps"] == 3 and j["value"] > 0 and j["gpu_launches"] > 0 and "workload" in j["config"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(j["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(j["e2e"]) and j["e2e"]["h2d_bytes_per_step"] > 0
    assert {"value", "unit", "cores", "kind", "sample"} <= set(j["cpu_baseline"]) and j["cpu_baseline"]["kind"] in ("reference", "port")
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"])


def test_bench_default_line_carries_the_three_headline_workloads():
    j = _line([])
    assert j["metric"] == "spmm_gflops" and j["scaling"] == "strong"
    assert set(j["workloads"]) == {"spmv", "mttkrp"}
    for rec in j["workloads"].values():
        assert rec["value"] > 0 and rec["gpu_launches"] > 0 and rec["e2e"]["h2d_bytes_per_step"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rec["roofline"])
        assert rec["cpu_baseline"]["kind"] in ("reference", "port")


def test_bench_reference_arm_contract():
    j = _line(["--impl", "reference"])
    assert j["impl"] == "reference" and j["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["sample"] == "the full configuration"
    assert set(j["workloads"]) == {"spmv", "mttkrp"}
    ours = _line(["--workload", "spmm", "--no-e2e", "--no-cpu"])
    assert ours["config"] == j["config"], "both arms must describe the same configuration"


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys, runpy\n"
            "sys.argv=['bench.py','--impl','reference','--small','--steps','1','--warmup','0','--workload','spmv']\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "print('MAPPED', 'libtaco_b200' in open('/proc/self/maps').read(), 'taco_b200' in sys.modules)\n") % os.path.join(ROOT, "bench.py")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert "MAPPED False False" in r.stdout, r.stdout[-1000:] + r.stderr[-1000:]
