"""GPU test of the batch preconditioner GEMM (mesh_deform_b200/csrc/batch_gemm_tc.cuh): the tcgen05 (3xTF32, accumulator in
tensor memory) kernel and the SIMT fp32 kernel against a double-precision CPU reference, via the test program
tests/cuda/batch_gemm_check (built by __graft_entry__.build() / make -C tests/cuda)."""
import os
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
CUDA_DIR = os.path.join(ROOT, "tests", "cuda")


@pytest.mark.parametrize("V,K", [(642, 256), (130, 70), (33, 5), (2048, 64)])
def test_tensor_core_batch_gemm_matches_reference(V, K):
    exe = os.path.join(CUDA_DIR, "batch_gemm_check")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", CUDA_DIR, "all"])
    out = subprocess.run([exe, str(V), str(K)], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0 and "GEMM CHECK OK" in out.stdout, out.stdout + out.stderr
