#!/bin/bash
# Round 2, GPU call g (1 GPU): sanity of the final bench path (autotune, one-rank sharder, two frames in flight) + GPU suite.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r02g_pytest_gpu.txt
timeout 400 python bench.py > gpurun_out/r02g_bench_n1.log 2>&1; tail -1 gpurun_out/r02g_bench_n1.log | tee gpurun_out/r02g_bench_n1.json | cut -c1-1500
for sc in triple_portal monoportal basics; do timeout 200 python bench.py --scene $sc --steps 100 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02g_${sc}_n1.json | cut -c1-200; done
timeout 300 python bench.py --scene mobius_monoportal --orbit 360 --steps 360 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02g_orbit_n1.json | cut -c1-1200
timeout 300 python bench.py --steps 200 --no-cpu-baseline --no-autotune --no-overlap 2>&1 | tail -1 | tee gpurun_out/r02g_bench_n1_plain.json | cut -c1-200
timeout 200 python tools/stream_roofline.py 3840x2160 50 2>&1 | tail -4 | tee gpurun_out/r02g_stream_roofline.txt | cut -c1-200
