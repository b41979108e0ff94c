#!/usr/bin/env python
"""What the box can deliver: aggregate device->host bandwidth with every GPU copying at once (one rank per GPU, torchrun).
The e2e leg of bench.py at N GPUs moves one RGBA8 frame per step from N GPUs to host memory; this is its roof.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/d2h_ceiling.py

Every rank copies a 64 MB device buffer into its own pinned host buffer (allocated next to its GPU), 40 times back to back,
timed with CUDA events between two barriers; rank 0 prints per-rank and aggregate GB/s, alone (ranks one at a time) and all
together, plus `nvidia-smi topo -m` (GPUs that share a PCIe switch share its uplink)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from portal_b200.distributed import gpu_numa_affinity  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    n = 64 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    with gpu_numa_affinity(local):
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        host.zero_()
    stream = torch.cuda.Stream()

    def run(reps=40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record()
            for _ in range(reps):
                host.copy_(dev, non_blocking=True)
            e1.record()
        stream.synchronize()
        return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    run(5)
    alone = []
    for r in range(world):
        dist.barrier()
        v = run() if r == rank else 0.0
        t = torch.tensor([v], dtype=torch.float64)
        dist.all_reduce(t)
        alone.append(round(float(t.item()), 1))
    dist.barrier()
    v = run()
    g = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(g, torch.tensor([v], dtype=torch.float64))
    if rank == 0:
        together = [round(float(x.item()), 1) for x in g]
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        print(json.dumps({"gpus": world, "d2h_gbs_one_rank_at_a_time": alone, "d2h_gbs_all_ranks_together": together,
                          "aggregate_gbs_together": round(sum(together), 1),
                          "rgba8_4k_frames_per_s_roof": round(sum(together) * 1e9 / (3840 * 2160 * 4), 1),
                          "mpixels_per_s_roof": round(sum(together) * 1e9 / 4 / 1e6, 1)}))
        print(topo)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
