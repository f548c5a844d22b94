#!/usr/bin/env bash
# Produces the ncu artefacts committed under profiles/ (run on the GPU box through gpurun):
#   1. launch list of a short bench run (per-launch device time, cold-cache & serialised: compare SHARES)
#   2. one --set full capture of the dominant kernel (pin_solve_kernel) late in the path
set -e
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:pin_solve -s 80 -c 1 -o gpurun_out/${TAG}_sweep -f \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_sweep_bench.log 2>&1 || true
ls -la gpurun_out
