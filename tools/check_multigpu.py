#!/usr/bin/env python
"""Run under torchrun (one rank per GPU): renders one frame sharded by cyclic row strips in both
device-side assembly modes (NCCL gather, and kernels storing straight into rank 0's frame over NVLink)
and through the host-delivery path (RGBA8 strips over every rank's PCIe link into one shared host frame),
and checks on rank 0 that each assembled frame is bit-identical to the single-GPU render."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from portal_b200.distributed import FrameSharder, HostFrameSharder  # noqa: E402
from portal_b200.renderer import SceneRenderer, load_scene_ir, load_textures  # noqa: E402


def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "portal_in_portal"
    w, h, depth = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (1920, 1080, 40)))
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sd = os.path.join(ROOT, "tests/golden/scenes")
    r = SceneRenderer(load_scene_ir(f"{sd}/{scene}.scene.json"), textures=load_textures(f"{sd}/{scene}.textures.npz"), device=local)
    r.render_depth = depth
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ref = r.render_host(w, h) if rank == 0 else None
    ok = True
    for mode in ("gather", "p2p"):
        sh = FrameSharder(r, w, h, rank, world, mode=mode)
        for i in range(5 if mode == "p2p" else 2):      # p2p: more frames than buffers, exercises the recycling flags
            sh.render(i, stream.cuda_stream)
            stream.synchronize()
            dist.barrier()
            if i < 4:
                sh.release(stream.cuda_stream)
        if rank == 0:
            got = np.empty((h, w, 4), dtype=np.float32)
            r._check(r._lib.pe_memcpy_d2h(r._ctx, got.ctypes.data, sh.frame_ptr, got.nbytes, None))
            same = np.array_equal(got.view(np.uint32), ref.view(np.uint32))
            print(f"multi-GPU {mode}: world {world} {scene} {w}x{h} -> bit-identical to single GPU: {same}", flush=True)
            ok = ok and same
        dist.barrier()
        sh.close()
    # host delivery: RGBA8 strips over every rank's PCIe link into one shared page-locked frame, 5 frames through a
    # ring of 3 with a moving camera; rank 0 checks each against its own single-GPU RGBA8 render
    cam = r.cam
    hfs = HostFrameSharder(r, w, h, rank, world)
    frames, prev, same = 5, None, True

    def check(f):
        hfs.complete(f)
        if rank == 0:
            got = hfs.wait_frame(f).copy()
            hfs.release(f)
            r.set_cam(cam["look_at"], cam["alpha"] + 0.05 * f, cam["beta"], cam["r"])
            return np.array_equal(got, r.render_host_rgba8(w, h))
        return True

    for f in range(frames):
        r.set_cam(cam["look_at"], cam["alpha"] + 0.05 * f, cam["beta"], cam["r"])
        hfs.submit()
        if prev is not None:
            same = check(prev) and same
        prev = f
    same = check(prev) and same
    dist.barrier()
    hfs.close()
    if rank == 0:
        print(f"multi-GPU host strips: world {world} {scene} {w}x{h} x {frames} frames -> identical to single GPU: {same}", flush=True)
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
