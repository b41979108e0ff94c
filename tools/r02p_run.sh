#!/bin/bash
# the final tuner's choice on the plane-heavy 4K config (the r02g line predates the w-aware candidate)
mkdir -p gpurun_out
timeout 70 python bench.py --scene triple_portal --steps 100 --no-cpu-baseline 2>/dev/null | tail -n 1 > gpurun_out/r02p_triple_portal_n1.json; cut -c1-160 gpurun_out/r02p_triple_portal_n1.json
exit 0
