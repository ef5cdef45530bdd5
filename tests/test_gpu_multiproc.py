"""One-process-per-GPU tests (the production layout): torch backend + DDP, uccl.collective-style
send/recv/allgather over the P2P engine, DeepEP Buffer over a torch process group.
Need >= 2 GPUs (skipped on single-GPU boxes; virtual-rank tests cover the kernels there)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("what", ["pg", "collective", "ep"])
def test_torchrun(what):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    port = 29600 + {"pg": 1, "collective": 2, "ep": 3}[what]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py"), what]
    env = dict(os.environ, UCCL_B200_TIMEOUT_MS="15000")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=500, env=env)
    sys.stdout.write(r.stdout[-3000:])
    sys.stderr.write(r.stderr[-3000:])
    assert r.returncode == 0
    assert f"mp_worker {what}: OK" in r.stdout
