"""Runs tests/multigpu_check.py under torchrun on 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_rows_and_restarts(exchange):
    """rows mode with the column sums exchanged by peer stores inside the EM tail
    kernel (default) and by ncclAllReduce (MXB_NO_P2P=1), plus restarts mode."""
    from mixemt_b200 import _lib
    if _lib.lib.mxb_device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    # torchrun pins OMP_NUM_THREADS to 1; the CPU oracle inside the check wants the cores
    env["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 2) // 2))
    if exchange == "nccl":
        env["MXB_NO_P2P"] = "1"
    else:
        env.pop("MXB_NO_P2P", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517" if exchange == "p2p" else "29518",
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert res.returncode == 0 and "MULTIGPU_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    assert ("exchange=%s" % exchange) in res.stdout, res.stdout[-2000:]
