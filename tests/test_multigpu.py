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


@pytest.mark.parametrize("world,size", [(2, (640, 360)), (3, (333, 100))])
def test_c_abi_sharder_protocol_on_one_gpu(world, size):
    """`pe_sharder_*` (include/portal_b200.h) -- owner / p2p / host modes, f32 and RGBA8 -- with `world` processes that all
    use GPU 0 (CUDA IPC works between processes on one device): every assembled frame is bit-identical to the single-GPU
    render.  Runs on the one-GPU box too, so the protocol that bench.py times at N > 1 is pixel-checked wherever the GPU
    suite runs; tools/check_sharder.py without --same-gpu is the same check across real GPUs.  (3 ranks at height 100:
    7 strips, ragged last strip, and a frame count that wraps both rings.)"""
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharder.py"), "portal_in_portal", str(size[0]), str(size[1]), "40",
           "--same-gpu"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    assert p.returncode == 0, out[-3000:]
    for what in ("owner f32", "owner rgba8", "owner overlapped f32", "p2p f32", "p2p rgba8", "host rgba8"):
        assert f"sharder {what}" in out, out[-3000:]
    assert "False" not in p.stdout


def test_sharder_across_gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharder.py"), "portal_in_portal", "1920", "1080", "40"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "False" not in p.stdout, (p.stdout + p.stderr)[-3000:]
