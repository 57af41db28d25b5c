"""Multi-GPU path on real GPUs (needs >= 2 on the box; the single-GPU round-end run skips it): the pipelined panel
exchange against the plain one, bit for bit, on every rank (tools/dist_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_pipelined_exchange_equals_plain_exchange():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2          # the configuration this check was verified on in round 1 (the fast pipelined path also on 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(ROOT, "tools", "dist_check.py"), "1024"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
