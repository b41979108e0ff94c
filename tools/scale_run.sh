mkdir -p gpurun_out
run() { n=$1; shift; timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%200)) bench.py --gpus $n --steps 100 --warmup 5 --no-cpu-baseline "$@" 2>/dev/null | tail -1; }
run 8 > gpurun_out/r01h_scale_n8.json
run 4 > gpurun_out/r01h_scale_n4.json
run 2 > gpurun_out/r01h_scale_n2.json
run 8 --mode gather > gpurun_out/r01h_scale_n8_gather.json
run 8 --scene mobius_monoportal --orbit 360 --steps 48 > gpurun_out/r01h_orbit_n8.json
timeout 60 python bench.py --steps 100 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r01h_scale_n1.json
for f in gpurun_out/r01h_*.json; do echo $f; cut -c1-160 $f; done
