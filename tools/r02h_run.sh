#!/bin/bash
# Round 2, GPU call h (1 GPU): w-aware matrix products off / on, all configs (kernel ms, same bits)
mkdir -p gpurun_out
for sc in portal_in_portal triple_portal monoportal basics; do timeout 200 python tools/sweep.py $sc '{"w_aware":[0,1]}' 20 2>&1 | tail -2; done | tee gpurun_out/r02h_sweep_w_aware.txt
timeout 300 python tools/sweep.py mobius_monoportal '{"w_aware":[0,1],"block_threads":[1024],"min_blocks":[1],"canon_rays":[0,1]}' 5 7680x4320x64 2>&1 | tail -4 | tee -a gpurun_out/r02h_sweep_w_aware.txt
timeout 300 python bench.py --steps 200 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02h_bench_n1.json | cut -c1-300
