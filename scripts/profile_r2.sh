#!/usr/bin/env bash
# Round-2 ncu artefacts (run on ONE GPU through gpurun); summaries are copied from gpurun_out/ into profiles/ by hand.
#   1. launch list of one bench path (per-launch device time; cold-cache & serialised: compare SHARES)
#   2. --set full capture of the fused sweep kernel late in the path + the per-launch algorithmic bytes of the same run
#   3. --set full capture of the tensor-core Gram-panel kernel inside a binomial (IRLS) path on a config-3 shard
set -x
TAG=${1:-r2}; SKIP=${2:-95}
mkdir -p gpurun_out
FLAGS="--steps 1 --warmup 0 --no-e2e --no-cpu --no-same-work"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $FLAGS > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pin_solve_batched -s $SKIP -c 1 -o gpurun_out/${TAG}_sweep -f \
    python bench.py $FLAGS --dump-launches gpurun_out/${TAG}_sweep_launches.json > gpurun_out/${TAG}_sweep_bench.log 2>&1
echo $SKIP > gpurun_out/${TAG}_sweep_skip.txt
N=125000 P=50000 L=40 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_gram_tc -s 70 -c 2 -o gpurun_out/${TAG}_gram_tc -f \
    python scripts/explore_c3.py > gpurun_out/${TAG}_gram_tc.log 2>&1
ls -la gpurun_out | tail -12
