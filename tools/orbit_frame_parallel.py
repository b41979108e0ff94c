#!/usr/bin/env python
"""Config 5 the other way ("for the orbit"): FRAME-level parallelism -- rank g renders whole frames
g, g + G, ... of the camera orbit into its own device buffers; no strips, no gather, no collective on the data path.
bench.py measures the row-tiled form BASELINE.json asks for; this prints the same quantity for the alternative so that
both can be reported.  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/orbit_frame_parallel.py --frames 360 --warmup 8

Timing: CUDA events on the launching stream around this rank's frames, barrier + synchronize on both sides, max over
ranks; value = all frames x pixels / that time.  Prints one JSON line on rank 0."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="mobius_monoportal")
    ap.add_argument("--width", type=int, default=7680)
    ap.add_argument("--height", type=int, default=4320)
    ap.add_argument("--depth", type=int, default=64)
    ap.add_argument("--frames", type=int, default=360, help="frames of one full turn, shared out over the ranks")
    ap.add_argument("--warmup", type=int, default=8, help="untimed frames per rank")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from portal_b200.renderer import SceneRenderer, load_scene_ir, load_textures

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scenes = os.path.join(ROOT, "tests", "golden", "scenes")
    r = SceneRenderer(load_scene_ir(os.path.join(scenes, f"{args.scene}.scene.json")),
                      textures=load_textures(os.path.join(scenes, f"{args.scene}.textures.npz")), device=local)
    r.render_depth = args.depth
    w, h = args.width, args.height
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sptr = stream.cuda_stream
    target = r.full_target(w, h)
    outs = [torch.empty((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(2)]   # 2 x 531 MB at 8K: larger than L2
    cam0 = dict(r.cam)
    mine = list(range(rank, args.frames, world))

    def frame(j, k):
        r.set_cam(cam0["look_at"], cam0["alpha"] + 2.0 * math.pi * k / args.frames, cam0["beta"], cam0["r"])
        r.draw_texture(target, outs[j & 1].data_ptr(), 0, sptr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for j in range(max(args.warmup, 3)):
        frame(j, mine[j % len(mine)])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = r.launch_count()
    barrier()
    e0.record(stream)
    for j, k in enumerate(mine):
        frame(j, k)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    n_launch = torch.tensor([r.launch_count() - l0], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_launch, op=dist.ReduceOp.SUM)
    if rank == 0:
        total_ms = float(ms.item())
        print(json.dumps({
            "metric": f"Mpixels/s @ {w}x{h} depth-{args.depth} {args.scene}", "value": round(w * h * args.frames / (total_ms * 1e-3) / 1e6, 2),
            "unit": "Mpixels/s", "n_gpus": world, "frames": args.frames, "ms_total": round(total_ms, 3), "higher_is_better": True,
            "scaling": "strong", "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.scene}.ron {w}x{h} depth {args.depth}, {args.frames}-frame camera orbit, aa 1",
                       "parallelism": f"{world} GPUs x whole frames (rank g renders frames g, g+{world}, ...), frames stay on their GPU, no collective"},
            "gpu_launches": int(n_launch.item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
