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


def test_bench_reference_arm_contract():
    j = _line(["--impl", "reference"])
    assert j["impl"] == "reference" and j["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["cores"] >= 1
