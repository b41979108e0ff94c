"""N > 1 host logic on CPU: world_size-2/3 gloo processes exchange synthetic strips through the same
layout + gather code the GPU path uses (portal_b200/distributed.py), and rank 0 checks that the
assembled frame is in row order.  (Rendering itself exists only on CUDA; on the GPU box
tests/test_parity_gpu.py::test_row_strips_reassemble_the_frame checks the kernels' side.)"""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from portal_b200 import distributed as D


def test_strip_layout_partitions_rows():
    for h, s, world in [(2160, 16, 8), (4320, 16, 8), (1080, 16, 2), (90, 16, 4), (17, 16, 8), (256, 8, 3)]:
        seen = []
        for rank in range(world):
            rows = D.local_rows(h, rank, world, s)
            assert len(rows) == D.strips_per_rank(h, world, s) * s
            seen += [y for y in rows if y >= 0]
            t = D.make_target(64, h, rank, world, s)
            assert t.n_strips == len(D.local_strips(h, rank, world, s)) and t.strip_first == rank and t.strip_step == world
        assert sorted(seen) == list(range(h))


def test_deinterleave_index_map():
    h, w, s, world = 50, 3, 16, 3
    spr = D.strips_per_rank(h, world, s)
    g = np.full((world, spr, s, w, 4), -1.0, dtype=np.float32)
    for rank in range(world):
        for lr, y in enumerate(D.local_rows(h, rank, world, s)):
            if y >= 0:
                g[rank].reshape(-1, w, 4)[lr] = y
    out = D.deinterleave_numpy(g, h, world, s)
    assert np.array_equal(out[:, 0, 0], np.arange(h, dtype=np.float32))


WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    from portal_b200 import distributed as D
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    h, w, s = {h}, {w}, 16
    spr = D.strips_per_rank(h, world, s)
    local = torch.full((spr, s, w, 4), -1.0)
    flat = local.view(-1, w, 4)
    for lr, y in enumerate(D.local_rows(h, rank, world, s)):
        if y >= 0:
            flat[lr] = torch.tensor([float(y), float(rank), 0.0, 1.0])
    for step in range(3):                         # the bench's per-frame sequence: render -> gather -> assemble
        gathered = D.gather_to_rank0(local, world, rank)
        if rank == 0:
            frame = D.deinterleave_numpy(gathered.numpy(), h, world, s)
            assert np.array_equal(frame[:, :, 0], np.broadcast_to(np.arange(h, dtype=np.float32)[:, None], (h, w)))
            owner = (np.arange(h) // s) % world
            assert np.array_equal(frame[:, 0, 1], owner.astype(np.float32))
        else:
            assert gathered is None
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    # the sharder's shared-memory segment name: every rank gets the SAME fresh name per sharder, a new one each time
    names = [D.shm_name("owner_f32", rank, world), D.shm_name("owner_f32", rank, world)]
    every = [None] * world
    dist.all_gather_object(every, names)
    assert all(n == every[0] for n in every) and names[0] != names[1], every
    assert D.shm_name("solo", 0, 1).endswith(f"_pid{{os.getpid()}}")          # a one-rank sharder inside a larger group needs no agreement
    dist.barrier()
    dist.destroy_process_group()
    print("RANK_OK", rank)
""")


@pytest.mark.parametrize("world,h", [(2, 2160), (3, 100)])
def test_gloo_gather_of_strips(world, h, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, h=h, w=8))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {rank}" in o, o[-2000:]


HOST_WORKER = textwrap.dedent("""
    import ctypes as C, os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r})
    from portal_b200 import distributed as D
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    h, w, s = {h}, {w}, 16

    class FakeLib:                       # stands in for the C ABI: "renders" strips whose bytes name (frame, row, rank)
        def __init__(self): self.frame = 0; self.tickets = 0
        def pe_host_register(self, ctx, p, n): return 0
        def pe_host_unregister(self, ctx, p): return 0
        def pe_sync(self, ctx): return 0
        def pe_wait_host(self, ctx, t): return 0
        def pe_submit_host_strips_rgba8(self, ctx, target, host_frame, ticket):
            t = target._obj
            for k in range(t.n_strips):
                row0 = (t.strip_first + k * t.strip_step) * t.strip_rows
                rows = max(0, min(t.strip_rows, t.height - row0))
                px = np.zeros((rows, t.width, 4), np.uint8)
                px[..., 0] = self.frame
                px[..., 1] = (np.arange(row0, row0 + rows) % 251)[:, None]
                px[..., 2] = rank
                px[..., 3] = 255
                C.memmove(host_frame + row0 * t.width * 4, px.ctypes.data, px.nbytes)
            self.frame += 1; self.tickets += 1
            ticket._obj.value = self.tickets
            return 0

    class FakeRenderer:
        def __init__(self): self._lib = FakeLib(); self._ctx = None
        def _check(self, rc): assert rc == 0
        def set_uniforms(self): pass

    hfs = D.HostFrameSharder(FakeRenderer(), w, h, rank, world, s)
    frames, prev = 8, None                       # more frames than ring slots: exercises the consumed counter

    def finish(f):
        hfs.complete(f)
        if rank == 0:
            fr = hfs.wait_frame(f)
            assert (fr[..., 0] == f).all() and (fr[..., 3] == 255).all()
            assert np.array_equal(fr[:, 0, 1], np.arange(h) % 251)
            assert np.array_equal(fr[:, 0, 2], (np.arange(h) // s) % world)
            hfs.release(f)

    for f in range(frames):
        assert hfs.submit() == f
        if prev is not None:
            finish(prev)
        prev = f
    finish(prev)
    dist.barrier()
    hfs.close()
    assert rank != 0 or not os.path.exists(hfs.path)
    dist.destroy_process_group()
    print("RANK_OK", rank)
""")


@pytest.mark.parametrize("world,h", [(2, 360), (3, 100)])
def test_gloo_host_frame_ring(world, h, tmp_path):
    """HostFrameSharder's shared-memory ring and counters with a stand-in for the C ABI (the real one needs GPUs:
    tools/check_multigpu.py)."""
    script = tmp_path / "worker.py"
    script.write_text(HOST_WORKER.format(root=ROOT, h=h, w=24))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"RANK_OK {rank}" in o, o[-2000:]
