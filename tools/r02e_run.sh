#!/bin/bash
# Round 2, GPU call e (2 GPUs): owner mode with two frames in flight, deeper host-delivery pipeline.
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
run 2 tools/check_sharder.py portal_in_portal 1920 1080 40 2>&1 | grep "^sharder" | tee gpurun_out/r02e_check_sharder_n2.txt
run 2 bench.py --gpus 2 --steps 200 --warmup 5 2>gpurun_out/r02e_n2.err | tail -1 | tee gpurun_out/r02e_scale_n2.json | cut -c1-200
run 2 bench.py --gpus 2 --steps 200 --warmup 5 --no-overlap --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02e_scale_n2_no_overlap.json | cut -c1-200
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r02e_scale_n1.json | cut -c1-200
tail -3 gpurun_out/r02e_n2.err
