"""Slab-partitioned solve on >= 2 GPUs of one box against the single-GPU solve (natural ordering).  Skipped on a
one-GPU box; run by hand with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("n", [2, 4])
def test_slab_solve_matches_single_gpu(n):
    if _ngpus() < n:
        pytest.skip("needs %d GPUs" % n)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + n), os.path.join(ROOT, "scripts", "slab_check.py"), "8"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "SLAB_CHECK PASS" in r.stdout, r.stdout[-4000:]
