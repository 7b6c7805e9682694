"""The record exchange folded into the measurement kernels (csrc/kernels.cuh: RecordSink +
exchange_signal_kernel) against an NCCL all-gather, byte for byte — needs >= 2 GPUs on the box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 8])
def test_peer_exchange_equals_nccl_all_gather(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "workers", "exchange_worker.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "EXCHANGE_OK world=%d" % world in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
