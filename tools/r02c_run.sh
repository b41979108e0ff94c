#!/bin/bash
# Round 2, GPU call c (2 GPUs): the C-ABI sharder across real GPUs, bench.py at N = 2 in every mode, parity blocks.
mkdir -p gpurun_out
run() { n=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) "$@"; }
echo "== sharder across GPUs"
run 2 tools/check_sharder.py portal_in_portal 1920 1080 40 2>&1 | grep -v "^W\|^\[W\|warn" | tail -8 | tee gpurun_out/r02c_check_sharder_n2.txt
run 2 tools/check_multigpu.py portal_in_portal 1920 1080 40 2>&1 | grep "multi-GPU" | tee gpurun_out/r02c_check_multigpu_n2.txt
echo "== bench N=2"
run 2 bench.py --gpus 2 --steps 200 --warmup 5 2>gpurun_out/r02c_n2_owner.err | tail -1 | tee gpurun_out/r02c_scale_n2.json | cut -c1-2500
run 2 bench.py --gpus 2 --steps 200 --warmup 5 --mode gather --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02c_scale_n2_gather.json | cut -c1-300
run 2 bench.py --gpus 2 --steps 200 --warmup 5 --mode p2p --tile-w 32 --no-assembled 2>/dev/null | tail -1 | tee gpurun_out/r02c_scale_n2_p2p_tile32.json | cut -c1-300
echo "== bench N=1 (same box)"
timeout 300 python bench.py --steps 200 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/r02c_scale_n1.json | cut -c1-300
echo "== reference arm under torchrun"
run 2 bench.py --impl reference --gpus 2 --steps 10 --warmup 2 2>/dev/null | tail -1 | tee gpurun_out/r02c_reference_n2.json | cut -c1-500
tail -5 gpurun_out/r02c_n2_owner.err
