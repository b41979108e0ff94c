"""Row-strip sharding of one frame over the GPUs of a box.

One process per GPU (torchrun); `torch.distributed` is the plumbing.  Pixels are independent, so
the path shards with no data-path collective except the assembly of the finished strips:

  * mode "gather": every rank renders its cyclic strips into a compact buffer, ONE NCCL gather
    brings them to rank 0, `pe_deinterleave_strips` (a streaming kernel) puts them in row order;
  * mode "p2p":    rank 0's frame is mapped into every other rank (CUDA IPC over NVLink) and each
    rank's render kernel stores its pixels straight into their final place while it is still
    computing the next ones -- the transfer is fused into the compute kernel, no collective, only a
    barrier at the end of the frame.

Strips are cyclic (rank r owns global strips r, r + G, ...) because cost per row is not uniform:
rows crossing portals run deeper.  The layout functions below are backend-agnostic (tested with gloo
on CPU); rendering itself only exists on CUDA.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import PeTarget

STRIP_ROWS = 16


def n_global_strips(height: int, strip_rows: int = STRIP_ROWS) -> int:
    return (height + strip_rows - 1) // strip_rows


def local_strips(height: int, rank: int, world: int, strip_rows: int = STRIP_ROWS):
    """Global strip indices owned by `rank`."""
    return list(range(rank, n_global_strips(height, strip_rows), world))


def strips_per_rank(height: int, world: int, strip_rows: int = STRIP_ROWS) -> int:
    """Strips every rank's compact buffer is padded to (rank 0 always has the most)."""
    return len(local_strips(height, 0, world, strip_rows))


def make_target(width: int, height: int, rank: int, world: int, strip_rows: int = STRIP_ROWS, full_frame: bool = False) -> PeTarget:
    return PeTarget(width, height, strip_rows, rank, world, len(local_strips(height, rank, world, strip_rows)), int(full_frame))


def local_rows(height: int, rank: int, world: int, strip_rows: int = STRIP_ROWS):
    """Global row index of every row of rank's compact buffer (-1 for padding rows)."""
    rows = []
    for g in local_strips(height, rank, world, strip_rows):
        rows += [y if y < height else -1 for y in range(g * strip_rows, (g + 1) * strip_rows)]
    pad = strips_per_rank(height, world, strip_rows) * strip_rows - len(rows)
    return rows + [-1] * pad


def gather_to_rank0(local, world: int, rank: int):
    """One gather of the equally sized compact buffers; returns [world, ...] on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local.unsqueeze(0)
    gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device) if rank == 0 else None
    dist.gather(local, list(gathered.unbind(0)) if rank == 0 else None, dst=0)
    return gathered


def deinterleave_numpy(gathered: np.ndarray, height: int, world: int, strip_rows: int = STRIP_ROWS) -> np.ndarray:
    """Host restatement of pe_deinterleave_strips' index map, for checking layouts in CPU tests."""
    out = np.empty((height,) + gathered.shape[3:], dtype=gathered.dtype)
    for rank in range(world):
        rows = local_rows(height, rank, world, strip_rows)
        flat = gathered[rank].reshape((-1,) + gathered.shape[3:])
        for lr, y in enumerate(rows):
            if y >= 0:
                out[y] = flat[lr]
    return out


class FrameSharder:
    """Per-rank state of a sharded render (CUDA only)."""

    def __init__(self, renderer, width: int, height: int, rank: int, world: int, mode: str = "gather", strip_rows: int = STRIP_ROWS):
        import torch
        import torch.distributed as dist
        assert mode in ("gather", "p2p")
        self.r, self.w, self.h, self.rank, self.world, self.mode, self.s = renderer, width, height, rank, world, mode, strip_rows
        self.spr = strips_per_rank(height, world, strip_rows)
        lib, ctx = renderer._lib, renderer._ctx
        self.frame_ptr = None      # rank 0: final frame (device pointer)
        self._peer_ptr = None
        if mode == "gather":
            self.target = make_target(width, height, rank, world, strip_rows, full_frame=False)
            self.local = [torch.zeros((self.spr, strip_rows, width, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
            if rank == 0:
                self.gathered = torch.empty((world, self.spr, strip_rows, width, 4), dtype=torch.float32, device="cuda")
                self.frame = torch.empty((height, width, 4), dtype=torch.float32, device="cuda")
                self.frame_ptr = self.frame.data_ptr()
        else:
            self.target = make_target(width, height, rank, world, strip_rows, full_frame=True)
            handle = torch.zeros(64, dtype=torch.uint8)
            if rank == 0:
                p = C.c_void_p()
                renderer._check(lib.pe_device_malloc(ctx, width * height * 16, C.byref(p)))
                self.frame_ptr = p.value
                hb = (C.c_uint8 * 64)()
                renderer._check(lib.pe_ipc_export(ctx, p, hb))
                handle = torch.tensor(list(hb), dtype=torch.uint8)
            if world > 1:
                h = handle.cuda()
                dist.broadcast(h, src=0)
                handle = h.cpu()
            if rank == 0:
                self.dst_ptr = self.frame_ptr
            else:
                hb = (C.c_uint8 * 64)(*handle.tolist())
                p = C.c_void_p()
                renderer._check(lib.pe_ipc_open(ctx, hb, C.byref(p)))
                self._peer_ptr = p.value
                self.dst_ptr = p.value

    def render(self, i: int, stream_ptr: int):
        """Render this rank's strips of frame i and assemble on rank 0 (asynchronous on the stream,
        except for the end-of-frame barrier in p2p mode which the caller issues)."""
        import torch.distributed as dist
        r = self.r
        if self.mode == "gather":
            out = self.local[i & 1]
            r.draw_texture(self.target, out.data_ptr(), 0, stream_ptr)
            if self.world > 1:
                dist.gather(out, list(self.gathered.unbind(0)) if self.rank == 0 else None, dst=0)
                if self.rank == 0:
                    r._check(r._lib.pe_deinterleave_strips(r._ctx, self.gathered.data_ptr(), self.frame_ptr, self.w, self.h,
                                                           self.s, self.world, self.spr, stream_ptr))
            elif self.rank == 0:
                r._check(r._lib.pe_deinterleave_strips(r._ctx, out.data_ptr(), self.frame_ptr, self.w, self.h, self.s, 1,
                                                       self.spr, stream_ptr))
        else:
            r.draw_texture(self.target, self.dst_ptr, 0, stream_ptr)

    def fence(self):
        """End-of-frame fence for p2p mode, stream-ordered and without a host synchronisation: a 4-byte
        all-reduce enqueued after the render kernel on every rank.  When it completes on rank 0's
        stream every rank's kernel -- and with it every remote store into rank 0's frame -- is done."""
        import torch
        import torch.distributed as dist
        if self.world > 1:
            if not hasattr(self, "_token"):
                self._token = torch.zeros(1, dtype=torch.int32, device="cuda")
            dist.all_reduce(self._token)

    def close(self):
        lib, ctx = self.r._lib, self.r._ctx
        if self._peer_ptr:
            lib.pe_ipc_close(ctx, self._peer_ptr)
            self._peer_ptr = None
        if self.mode == "p2p" and self.rank == 0 and self.frame_ptr:
            lib.pe_device_free(ctx, self.frame_ptr)
            self.frame_ptr = None
