#!/usr/bin/env python
"""Which part of the host-delivery path costs bandwidth?  One GPU: a 16.6 MB device buffer (half a 4K RGBA8 frame) copied to
host memory 60 times, as (a) one contiguous copy into cudaHostAlloc'ed memory, (b) the same into a cudaHostRegister'ed POSIX
shared-memory mapping, (c) / (d) as 68 strips of 245 760 B with a destination pitch of twice that (the sharder's 2-rank
layout) into either, (e) = (d) while a render-like kernel keeps the SMs busy.  GB/s by CUDA events."""
import ctypes as C
import json
import mmap
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
D2H = 2


def main():
    torch.cuda.set_device(0)
    strip, n = 245760, 68
    nbytes = strip * n
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    pinned = torch.empty(2 * nbytes, dtype=torch.uint8).pin_memory()
    path = f"/dev/shm/pe_d2h_variants_{os.getpid()}"
    fd = os.open(path, os.O_CREAT | os.O_RDWR, 0o600)
    os.ftruncate(fd, 2 * nbytes)
    mm = mmap.mmap(fd, 2 * nbytes)
    os.close(fd)
    import numpy as np
    arr = np.frombuffer(mm, dtype=np.uint8)
    arr[:] = 0
    assert rt.cudaHostRegister(arr.ctypes.data, 2 * nbytes, 1) == 0
    s = torch.cuda.Stream()
    busy = torch.cuda.Stream()
    a = torch.randn(4096, 4096, device="cuda")
    out = {}

    def timed(fn, reps=60, with_busy=False):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if with_busy:
            with torch.cuda.stream(busy):
                for _ in range(40):
                    torch.mm(a, a)
        e0.record(s)
        for _ in range(reps):
            fn()
        e1.record(s)
        torch.cuda.synchronize()
        return round(nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    sp = s.cuda_stream
    out["contiguous -> cudaHostAlloc"] = timed(lambda: rt.cudaMemcpyAsync(pinned.data_ptr(), dev.data_ptr(), nbytes, D2H, sp))
    out["contiguous -> registered shm"] = timed(lambda: rt.cudaMemcpyAsync(arr.ctypes.data, dev.data_ptr(), nbytes, D2H, sp))
    out["68 strips, pitch x2 -> cudaHostAlloc"] = timed(lambda: rt.cudaMemcpy2DAsync(pinned.data_ptr(), 2 * strip, dev.data_ptr(), strip, strip, n, D2H, sp))
    out["68 strips, pitch x2 -> registered shm"] = timed(lambda: rt.cudaMemcpy2DAsync(arr.ctypes.data, 2 * strip, dev.data_ptr(), strip, strip, n, D2H, sp))
    out["68 strips as 68 contiguous copies -> registered shm"] = timed(
        lambda: [rt.cudaMemcpyAsync(arr.ctypes.data + 2 * strip * k, dev.data_ptr() + strip * k, strip, D2H, sp) for k in range(n)])
    out["68 strips, pitch x2 -> registered shm, SMs busy"] = timed(
        lambda: rt.cudaMemcpy2DAsync(arr.ctypes.data, 2 * strip, dev.data_ptr(), strip, strip, n, D2H, sp), with_busy=True)
    out["contiguous -> cudaHostAlloc, SMs busy"] = timed(lambda: rt.cudaMemcpyAsync(pinned.data_ptr(), dev.data_ptr(), nbytes, D2H, sp), with_busy=True)
    print(json.dumps(out))
    os.unlink(path)


if __name__ == "__main__":
    main()
