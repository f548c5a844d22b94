#!/usr/bin/env bash
# usage: scripts/run_ncu.sh <tag> [skip] -- captures one pin_solve_kernel launch with --set full
set -e
TAG=${1:-prof}; SKIP=${2:-40}
mkdir -p gpurun_out
N=${N:-200000} P=${P:-2000} L=${L:-60} ncu --set full --clock-control none --import-source on -k regex:pin_solve -s $SKIP -c 1 \
  -o gpurun_out/$TAG -f python scripts/explore_c2.py > gpurun_out/$TAG.log 2>&1
tail -5 gpurun_out/$TAG.log
