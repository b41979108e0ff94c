#!/bin/bash
# the autotune guard + pe_frames_differ on a GPU (the round's last GPU seconds)
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "frames_differ or autotune_keeps" 2>&1 | tail -n 5 | tee gpurun_out/r02o_pytest_guard.txt
timeout 60 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>gpurun_out/r02o_bench.err | tail -n 1 > gpurun_out/r02o_bench_n1.json; cut -c1-200 gpurun_out/r02o_bench_n1.json
exit 0
