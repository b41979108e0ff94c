#!/bin/bash
# Round 2, GPU call j (2 GPUs): host delivery with NUMA placement through NVML, against the box's D2H ceiling.
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
run 2 tools/d2h_ceiling.py 2>/dev/null | head -1 | tee gpurun_out/r02j_d2h_ceiling_n2.txt | cut -c1-300
PORTAL_B200_DEBUG=1 run 2 bench.py --gpus 2 --steps 200 --warmup 5 --no-assembled 2>gpurun_out/r02j_n2.err | tail -1 | tee gpurun_out/r02j_scale_n2.json | cut -c1-200
grep pe_sharder gpurun_out/r02j_n2.err | head -4
run 2 tools/check_sharder.py portal_in_portal 1920 1080 40 2>&1 | grep "^sharder" | tee gpurun_out/r02j_check_sharder_n2.txt
