#!/bin/bash
# First GPU call of the next round: everything that changed after round 1's GPU budget ran out, in one go (1 GPU, ~15 min).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_run.sh'
# Outputs land in gpurun_out/r02a_*; nothing here is a bench value except the plain bench.py line.
mkdir -p gpurun_out
echo "== full GPU suite"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02a_pytest_gpu.txt
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== headline"
timeout 300 python bench.py > gpurun_out/r02a_bench_n1.log 2>&1; tail -1 gpurun_out/r02a_bench_n1.log | tee gpurun_out/r02a_bench_n1.json | cut -c1-600
echo "== candidates of DESIGN.md section 10 (kernel ms, same frame)"
timeout 200 python tools/sweep.py portal_in_portal '{"uniforms_in_smem":[0,1]}' 20 2>&1 | tail -4 | tee gpurun_out/r02a_sweep_smem.txt
timeout 200 python tools/sweep.py basics '{"block_threads":[128,256,512],"min_blocks":[2,4]}' 200 256x256x4 2>&1 | tail -8 | tee gpurun_out/r02a_sweep_small_target.txt
echo "== launch list of the bench command"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02a_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -3 gpurun_out/r02a_launches_bench.csv
# on an 8-GPU call (gpurun --gpus 8), the orbit both ways:
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/orbit_frame_parallel.py > gpurun_out/r02a_orbit_frames_n8.json
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --scene mobius_monoportal --orbit 360 --steps 48 --no-cpu-baseline > gpurun_out/r02a_orbit_rows_n8.json
