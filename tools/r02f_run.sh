#!/bin/bash
# Round 2, GPU call f (1 GPU): the current kernel (canonical rays, NVRTC 12.9) -- GPU suite, ncu --set full on all five
# configs (for tools/sass_attribution.py), sweeps, streaming kernels, every config's bench line.
mkdir -p gpurun_out
echo "== GPU suite"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r02f_pytest_gpu.txt
echo "== canonical rays off / on, all configs (kernel ms, same bits)"
for sc in portal_in_portal triple_portal monoportal basics; do timeout 200 python tools/sweep.py $sc '{"canon_rays":[0,1]}' 20 2>&1 | tail -2; done | tee gpurun_out/r02f_sweep_canon.txt
timeout 300 python tools/sweep.py mobius_monoportal '{"canon_rays":[0,1]}' 5 7680x4320x64 2>&1 | tail -2 | tee -a gpurun_out/r02f_sweep_canon.txt
echo "== register cap / block size with canonical rays"
timeout 400 python tools/sweep.py portal_in_portal '{"block_threads":[256,512],"min_blocks":[2,3,4]}' 20 2>&1 | tail -6 | tee gpurun_out/r02f_sweep_blocks.txt
timeout 400 python tools/sweep.py portal_in_portal,triple_portal '{"persistent":[0,1]}' 20 2>&1 | tail -4 | tee gpurun_out/r02f_sweep_persistent.txt
echo "== mobius: code size vs instruction cache"
timeout 400 python tools/sweep.py mobius_monoportal '{"unroll_loops":[1,0],"block_threads":[512,1024],"min_blocks":[1]}' 5 2>&1 | tail -4 | tee gpurun_out/r02f_sweep_mobius.txt
timeout 400 python tools/sweep.py mobius_monoportal '{"unroll_loops":[1,0],"block_threads":[512],"min_blocks":[2]}' 5 2>&1 | tail -2 | tee -a gpurun_out/r02f_sweep_mobius.txt
echo "== ncu --set full, one launch per config scene at its BASELINE size"
cfg() { case $1 in portal_in_portal|triple_portal) echo 3840x2160x40;; monoportal) echo 1920x1080x20;; mobius_monoportal) echo 7680x4320x64;; basics) echo 256x256x4;; esac; }
for sc in portal_in_portal triple_portal monoportal mobius_monoportal basics; do
  timeout 300 ncu --set full --clock-control none --import-source on -k pe_render_kernel -s 3 -c 1 -f -o gpurun_out/r02f_ncu_$sc \
      python tools/sweep.py $sc '{}' 1 $(cfg $sc) > gpurun_out/r02f_ncu_$sc.log 2>&1
  tail -1 gpurun_out/r02f_ncu_$sc.log | cut -c1-200
done
echo "== streaming kernels"
timeout 200 python tools/stream_roofline.py 3840x2160 50 2>&1 | tail -6 | tee gpurun_out/r02f_stream_roofline.txt
echo "== bench lines"
timeout 400 python bench.py > gpurun_out/r02f_bench_n1.log 2>&1; tail -1 gpurun_out/r02f_bench_n1.log | tee gpurun_out/r02f_bench_n1.json | cut -c1-400
for sc in triple_portal monoportal basics; do timeout 200 python bench.py --scene $sc --steps 100 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02f_${sc}_n1.json | cut -c1-200; done
timeout 300 python bench.py --scene mobius_monoportal --orbit 360 --steps 360 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02f_orbit_n1.json | cut -c1-200
echo "== launch list of the bench command"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -2 gpurun_out/r02f_launches_bench.csv
