#!/bin/bash
# Round 2, last GPU call (1 GPU), on the frozen tree: GPU suite, smoke, both bench arms, one ncu --set full of the
# headline kernel as autotune ships it (defaults: 512-thread blocks, canonical rays, w-aware products), launch list.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 | tee gpurun_out/r02n_pytest_gpu.txt
timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2 | tee gpurun_out/r02n_smoke.txt
timeout 200 python bench.py > gpurun_out/r02n_bench_n1.log 2>&1; tail -n 1 gpurun_out/r02n_bench_n1.log | tee gpurun_out/r02n_bench_n1.json | cut -c1-300
timeout 150 python bench.py --impl reference 2>/dev/null | tail -n 1 | tee gpurun_out/r02n_bench_reference.json | cut -c1-300
timeout 150 ncu --set full --clock-control none --import-source on -k pe_render_kernel -s 3 -c 1 -f -o gpurun_out/r02n_ncu_portal_in_portal \
    python tools/sweep.py portal_in_portal '{}' 1 3840x2160x40 > gpurun_out/r02n_ncu_portal_in_portal.log 2>&1
tail -n 1 gpurun_out/r02n_ncu_portal_in_portal.log | cut -c1-200
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02n_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
tail -n 2 gpurun_out/r02n_launches_bench.csv | cut -c1-300
exit 0
