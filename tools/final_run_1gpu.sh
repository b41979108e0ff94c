mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "rgba8 or golden" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_final3.log 2>&1; tail -1 gpurun_out/bench_final3.log | cut -c1-400
for sc in monoportal triple_portal; do timeout 100 python bench.py --scene $sc --steps 100 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r01i_${sc}_n1.json; cut -c1-200 gpurun_out/r01i_${sc}_n1.json; done
timeout 150 python bench.py --scene mobius_monoportal --orbit 360 --steps 36 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r01i_orbit_n1.json; cut -c1-200 gpurun_out/r01i_orbit_n1.json
timeout 100 python bench.py --scene basics --steps 200 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r01i_basics_n1.json; cut -c1-200 gpurun_out/r01i_basics_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
