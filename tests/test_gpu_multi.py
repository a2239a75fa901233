"""Multi-GPU parity on a box with at least two GPUs (skipped otherwise): tile-parallel rendering straight into the
presenting GPU's image over peer memory (fr_ipc_*) is bit-identical to the single-GPU image.  The host-side partition
logic is covered without GPUs by tests/test_multigpu_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_tile_parallel_over_peer_memory_is_bit_identical():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on one box")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_tiles_peer.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("bit-identical") == 2 and "MISMATCH" not in r.stdout
