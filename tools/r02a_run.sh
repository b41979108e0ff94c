#!/bin/bash
# Round 2, first GPU call (1 GPU): the tree as round 1 left it + the persistent-AA fix.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02a_run.sh'
# Outputs: gpurun_out/r02a_*.  Only the plain bench.py line is a bench value; everything under ncu is for counters.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02a_pytest_gpu.txt
echo "== headline bench"
timeout 300 python bench.py > gpurun_out/r02a_bench_n1.log 2>&1; tail -1 gpurun_out/r02a_bench_n1.log | tee gpurun_out/r02a_bench_n1.json | cut -c1-400
echo "== streaming kernels vs the HBM roof"
timeout 200 python tools/stream_roofline.py 3840x2160 50 2>&1 | tail -6 | tee gpurun_out/r02a_stream_roofline.txt
echo "== uniform block: constant bank vs shared-memory image, all five configs"
for sc in portal_in_portal triple_portal monoportal basics; do
  timeout 200 python tools/sweep.py $sc '{"uniforms_in_smem":[0,1,2]}' 20 2>&1 | tail -3
done | tee gpurun_out/r02a_sweep_smem.txt
timeout 200 python tools/sweep.py mobius_monoportal '{"uniforms_in_smem":[0,1,2]}' 5 7680x4320x64 2>&1 | tail -3 | tee -a gpurun_out/r02a_sweep_smem.txt
echo "== warp tile shape 8x4 / 16x2 / 32x1 (local stores)"
for sc in portal_in_portal triple_portal monoportal; do
  timeout 200 python tools/sweep.py $sc '{"tile_w":[8,16,32]}' 20 2>&1 | tail -3
done | tee gpurun_out/r02a_sweep_tile.txt
echo "== ncu --set full, one launch per config scene at its BASELINE size"
cfg() { case $1 in portal_in_portal|triple_portal) echo 3840x2160x40;; monoportal) echo 1920x1080x20;; mobius_monoportal) echo 7680x4320x64;; basics) echo 256x256x4;; esac; }
for sc in portal_in_portal triple_portal monoportal mobius_monoportal basics; do
  timeout 300 ncu --set full --clock-control none --import-source on -k pe_render_kernel -s 3 -c 1 -f -o gpurun_out/r02a_ncu_$sc \
      python tools/sweep.py $sc '{}' 1 $(cfg $sc) > gpurun_out/r02a_ncu_$sc.log 2>&1
  tail -1 gpurun_out/r02a_ncu_$sc.log | cut -c1-200
done
echo "== launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02a_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r02a_launches_bench.csv
ls -la gpurun_out | head -40
