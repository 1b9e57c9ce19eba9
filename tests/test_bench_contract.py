"""CPU tests of bench.py's contract: the reference arm (the CPU oracle alone) runs without a GPU and prints one JSON line with
the keys the driver reads; the GPU arm refuses to run without a device instead of falling back to anything."""
import json
import os
import subprocess
import sys

from conftest import ROOT, _cuda_available

BENCH = os.path.join(ROOT, "bench.py")


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--nu", "8", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "arap_iterations_per_sec_1M_verts" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["vertices"] == 642 and d["value"] > 0


def test_gpu_arm_needs_a_device():
    if _cuda_available():
        return
    out = subprocess.run([sys.executable, BENCH, "--nu", "8", "--steps", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stdout + out.stderr)
