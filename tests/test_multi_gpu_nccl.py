"""Partitioned mode over NCCL on 2 GPUs (skipped on boxes with fewer). Launches torchrun on tests/tools/dist_partitioned_check.py."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("transport", ["nccl", "peer"])
def test_partitioned_mesh_over_nccl_two_gpus(transport):
    """One process per GPU: halo exchange + all-reduced CG + the row-partitioned global multigrid hierarchy, over NCCL and over
    the peer-memory transport (direct stores into the neighbour's mailbox through CUDA IPC); parity against the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    env = dict(os.environ, ARAP_DIST_TRANSPORT=transport)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if transport == "nccl" else "29518", os.path.join(ROOT, "tests", "tools", "dist_partitioned_check.py"), "96", "64", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in out.stdout.splitlines() if l.startswith("PARTITIONED ")]
    assert out.returncode == 0 and lines, out.stdout[-2000:] + out.stderr[-2000:]
    res = json.loads(lines[-1][len("PARTITIONED "):])
    assert res["ok"], res
    assert res["transport"].startswith("peer" if transport == "peer" else "nccl") and res["cg_graph"] == 1
