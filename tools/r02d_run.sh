#!/bin/bash
# Round 2, GPU call d (8 GPUs): the 1 -> 8 curve of bench.py (owner mode + assembled rates + host delivery), config 5 as
# specified (7680x4320 depth 64, 360-frame orbit, 8 GPUs), the sharder check across 8 GPUs.
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
echo "== sharder across 8 GPUs"
run 8 tools/check_sharder.py portal_in_portal 1920 1080 40 2>&1 | grep "^sharder" | tee gpurun_out/r02d_check_sharder_n8.txt
echo "== scale"
for n in 8 4 2; do
  run $n bench.py --gpus $n --steps 200 --warmup 5 2>gpurun_out/r02d_n$n.err | tail -1 | tee gpurun_out/r02d_scale_n$n.json | cut -c1-200
done
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r02d_scale_n1.json | cut -c1-200
echo "== config 5: 8K orbit, 360 frames, 8 GPUs (rows) and 1 GPU"
run 8 bench.py --gpus 8 --scene mobius_monoportal --orbit 360 --steps 360 --warmup 5 2>gpurun_out/r02d_orbit.err | tail -1 | tee gpurun_out/r02d_orbit_n8.json | cut -c1-200
echo "== gather mode at 8 (north_star's single NCCL gather)"
run 8 bench.py --gpus 8 --steps 200 --warmup 5 --mode gather --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02d_scale_n8_gather.json | cut -c1-200
run 8 bench.py --gpus 8 --steps 200 --warmup 5 --mode p2p --tile-w 32 --format rgba8 --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02d_scale_n8_p2p_rgba8_tile32.json | cut -c1-200
tail -3 gpurun_out/r02d_n8.err gpurun_out/r02d_orbit.err
