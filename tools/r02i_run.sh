#!/bin/bash
# Round 2, GPU call i (8 GPUs): the 1 -> 8 curve with the final bench path (autotune, two frames in flight, deeper host
# pipeline), the box's D2H ceiling, config 5 as specified.
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
echo "== what the box can deliver to host memory"
run 8 tools/d2h_ceiling.py 2>/dev/null | tee gpurun_out/r02i_d2h_ceiling_n8.txt | head -3 | cut -c1-600
run 2 tools/d2h_ceiling.py 2>/dev/null | head -1 | tee gpurun_out/r02i_d2h_ceiling_n2.txt | cut -c1-400
echo "== scale"
for n in 8 4 2; do
  run $n bench.py --gpus $n --steps 200 --warmup 5 2>gpurun_out/r02i_n$n.err | tail -1 | tee gpurun_out/r02i_scale_n$n.json | cut -c1-200
done
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r02i_scale_n1.json | cut -c1-200
echo "== config 5: 8K orbit, 360 frames, 8 GPUs"
run 8 bench.py --gpus 8 --scene mobius_monoportal --orbit 360 --steps 360 --warmup 5 2>gpurun_out/r02i_orbit.err | tail -1 | tee gpurun_out/r02i_orbit_n8.json | cut -c1-200
echo "== for the record at 8: NCCL gather; one frame in flight"
run 8 bench.py --gpus 8 --steps 200 --warmup 5 --mode gather --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02i_scale_n8_gather.json | cut -c1-200
run 8 bench.py --gpus 8 --steps 200 --warmup 5 --no-overlap --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02i_scale_n8_no_overlap.json | cut -c1-200
tail -n 3 gpurun_out/r02i_n8.err
