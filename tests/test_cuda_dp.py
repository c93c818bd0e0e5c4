"""GPU (2+ devices): the data-parallel step through NCCL - the bucketed in-place all-reduce of the flat gradient arena,
the scaler's 1/world averaging and FlatAdamW - against a single-process recomputation of every rank's gradient
(scripts/gpu_check_dp.py, launched as one process per GPU with torch.distributed.run).  Skipped on a 1-GPU box; the
host-side logic of the same path runs at world size 2 over gloo in tests/test_parallel_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 CUDA devices (NCCL data-parallel check)")
def test_nccl_data_parallel_step_matches_local_recomputation():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "gpu_check_dp.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0 and "DP CHECK OK" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
