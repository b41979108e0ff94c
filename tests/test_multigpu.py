"""Multi-GPU assembly on real GPUs (skipped unless the box has >= 2): spawns torchrun over
tools/check_multigpu.py, which compares both assembly modes bit-for-bit with the single-GPU frame."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_sharded_frame_equals_single_gpu_frame():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_multigpu.py"), "portal_in_portal", "1920", "1080", "40"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, (p.stdout + p.stderr)[-3000:]
    assert "gather" in p.stdout and "p2p" in p.stdout and "host strips" in p.stdout and "False" not in p.stdout
