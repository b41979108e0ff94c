#!/bin/bash
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
PORTAL_B200_DEBUG=1 run 2 bench.py --gpus 2 --steps 200 --warmup 5 --no-assembled 2>gpurun_out/r02l_n2.err | tail -1 | tee gpurun_out/r02l_scale_n2.json | cut -c1-100
grep "bench\]" gpurun_out/r02l_n2.err
timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r02l_scale_n1.json | cut -c1-100
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_capi.py -m gpu -x -q 2>&1 | tail -2
