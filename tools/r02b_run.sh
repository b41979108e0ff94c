#!/bin/bash
# Round 2, GPU call b (1 GPU): the whole GPU suite on the tree with the C-ABI sharder, NVRTC pinned to the toolkit's, the
# new bench.py (parity block, .ron front-end on the e2e leg), config scenes at full size through both front-ends.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02b_pytest_gpu.txt
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== headline bench (ours, then the reference arm)"
timeout 400 python bench.py > gpurun_out/r02b_bench_n1.log 2>&1; tail -1 gpurun_out/r02b_bench_n1.log | tee gpurun_out/r02b_bench_n1.json | cut -c1-3000
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/r02b_bench_reference.json | cut -c1-600
echo "== other configs"
for sc in triple_portal monoportal basics; do timeout 200 python bench.py --scene $sc --steps 100 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02b_${sc}_n1.json | cut -c1-400; done
timeout 300 python bench.py --scene mobius_monoportal --orbit 360 --steps 360 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02b_orbit_n1.json | cut -c1-1500
echo "== streaming kernels"
timeout 200 python tools/stream_roofline.py 3840x2160 50 2>&1 | tail -6 | tee gpurun_out/r02b_stream_roofline.txt
