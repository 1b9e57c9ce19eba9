"""CPU test (world_size 2, gloo) of the host-side plumbing of `bench.py --gpus N`: the multi-GPU arms' collectives, the
all-ranks-agree error handling, the scaling arithmetic and the JSON assembly run with stand-ins for the engine
(tests/tools/bench_multi_fake.py). The oracle leg of the arm `partitioned_vs_oracle` is real (a 12 x 12 grid)."""
import json
import os
import socket
import subprocess
import sys

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(part_nx, weak, oracle_nx, budget):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "tools", "bench_multi_fake.py"), str(part_nx), str(weak), str(oracle_nx), str(budget)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("LINE ")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0][5:])["multi_gpu"]


def test_multi_gpu_arms_plumbing_two_ranks():
    m = _run(24, 200, 12, 500)
    assert m["world"] == 2
    for arm in ("partitioned_strong_strips", "partitioned_strong_blocks"):
        a = m[arm]
        assert "error" not in a and "skipped" not in a, a
        assert a["exchanges_per_cg_iteration"] == 8 and a["allreduces_per_cg_iteration"] == 2
        assert a["parity_vs_single_gpu"]["max_dp_over_bbox_diag"] == 0.0          # the stand-ins agree exactly
        assert abs(a["parity_vs_single_gpu"]["rel_energy_diff"]) < 1e-12
        assert abs(a["strong_scaling_efficiency"] - a["speedup_vs_single_gpu"] / 2) < 1e-12
        assert "positions" not in a and "limiter" in a
    w = m["partitioned_weak"]
    assert "weak_scaling_efficiency" in w and w["workload"].startswith("20 x 20")          # sqrt(200 * 2) = 20
    assert m["weak_single_gpu_reference"]["iterations_per_s"] > 0 if "weak_single_gpu_reference" in m else True
    assert "parity" in m["partitioned_vs_oracle"]                                        # the real CPU oracle ran on rank 0
    b = m["batch_spheres"]
    assert b["member_iterations_per_s"] > 0 and "strong_scaling_efficiency" in b
    json.dumps(m)                                                                        # nothing unserialisable left


def test_multi_gpu_arms_respect_the_time_budget():
    m = _run(24, 200, 12, 0)                       # budget 0 s: every arm is skipped, on every rank alike (no deadlock)
    assert all("skipped" in m[k] for k in ("partitioned_strong_strips", "partitioned_strong_blocks", "batch_spheres"))
