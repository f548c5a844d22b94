#!/usr/bin/env bash
# quick launch list of a reduced config-5 path (per-kernel device time)
set -e
TAG=${1:-snpq}
mkdir -p gpurun_out
export N=${N:-500000} P=${P:-20000} L=${L:-8} REPS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/explore_c5.py > gpurun_out/${TAG}_launches.log 2>&1 || true
tail -4 gpurun_out/${TAG}_launches.log
