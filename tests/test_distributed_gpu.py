"""Multi-GPU path on real GPUs (needs >= 2 on the box; the single-GPU round-end run skips it, bench.py --gpus N carries the same
checks in its `parity` object): the C ABI entry gemmul8_b200_pgemm with both transports (NCCL, copy engines over CUDA IPC) and
round 1's Python orchestration, each against the plain exchange and the unpartitioned product, bit for bit, on every rank
(tools/dist_check.py under torchrun)."""
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
    world = 4 if n >= 4 else 2          # 2 x 2 exercises the B-panel exchange too (P > 1)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(ROOT, "tools", "dist_check.py"), "1024", "all"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DIST OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
