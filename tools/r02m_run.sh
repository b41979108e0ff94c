#!/bin/bash
# 2 GPUs, bench only: the e2e loop no longer times the checker's private copy of the last frame (bench.py)
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
PORTAL_B200_DEBUG=1 run 2 bench.py --gpus 2 --steps 200 --warmup 5 --no-assembled 2>gpurun_out/r02m_n2.err | tail -n 1 | tee gpurun_out/r02m_scale_n2.json | cut -c1-100
grep "bench\]" gpurun_out/r02m_n2.err
PORTAL_B200_DEBUG=1 run 2 bench.py --gpus 2 --steps 20 --warmup 3 --no-assembled 2>gpurun_out/r02m_n2_k20.err | tail -n 1 | tee gpurun_out/r02m_scale_n2_k20.json | cut -c1-100
exit 0
