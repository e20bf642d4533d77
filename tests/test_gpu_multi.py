"""Multi-GPU checks that need >= 2 devices on the box (skipped otherwise): launched through torchrun exactly as
bench.py is, one rank per GPU over NCCL."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, script, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29000 + os.getpid() % 2000), script, *args]
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_dp_gradients_equal_global_batch_gradients():
    """BASELINE config 5: after the flat-bucket NCCL all-reduce every rank holds the gradient of the global batch
    (tests/dp_equivalence_check.py asserts <= 1e-4 relative per parameter)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, os.path.join(ROOT, "tests", "dp_equivalence_check.py"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "dp_equivalence world=2" in r.stdout


def test_sharded_forward_equals_single_gpu_forward():
    """T8: the same global batch on 1 vs 2 GPUs, Philox keyed by the global scene index -> bit-identical scenes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _torchrun(2, os.path.join(ROOT, "tests", "shard_invariance_check.py"))
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "shard_invariance world=2 ok" in r.stdout
