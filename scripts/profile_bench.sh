#!/usr/bin/env bash
# Produces the ncu artefacts summarised under profiles/ (run on the GPU box through gpurun):
#   1. the bench line of the default run
#   2. launch list of a short bench run (per-launch device time, cold-cache & serialised: compare SHARES)
#   3. one --set full capture of the dominant kernel late in the path, plus the per-launch algorithmic bytes of the same run
#      (bench.py --dump-launches) so that the captured launch's DRAM traffic can be set against its algorithmic bytes
set -e
TAG=${1:-r1}; SKIP=${2:-95}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err || true
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:pin_solve_batched -s $SKIP -c 1 -o gpurun_out/${TAG}_sweep -f \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu --dump-launches gpurun_out/${TAG}_sweep_launches.json > gpurun_out/${TAG}_sweep_bench.log 2>&1 || true
echo $SKIP > gpurun_out/${TAG}_sweep_skip.txt
ls -la gpurun_out
