#!/usr/bin/env bash
# ncu artefacts of the covariance-method solver (one GPU, through gpurun): launch list of one 100-lambda path and a --set full
# capture of the cluster kernel late in the path.  Summaries are copied from gpurun_out/ into profiles/ by hand.
set -x
TAG=${1:-r2_cov}; SKIP=${2:-80}
mkdir -p gpurun_out
COV_CPU=0 COV_REPS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/bench_cov.py 4000 10 10000 > gpurun_out/${TAG}_launches.log 2>&1
COV_CPU=0 COV_REPS=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:cov_pin_kernel -s $SKIP -c 1 -o gpurun_out/${TAG}_kernel -f \
    python scripts/bench_cov.py 4000 10 10000 > gpurun_out/${TAG}_kernel.log 2>&1
ncu -i gpurun_out/${TAG}_kernel.ncu-rep --page raw --csv > gpurun_out/${TAG}_kernel_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_kernel.ncu-rep --page details > gpurun_out/${TAG}_kernel_details.txt 2>/dev/null
COV_PROF=1 COV_CPU=1 timeout 300 python scripts/bench_cov.py 4000 10 10000 > gpurun_out/${TAG}_bench.log 2>&1
COV_PROF=1 COV_CPU=0 timeout 300 python scripts/bench_cov.py 8000 10 10000 >> gpurun_out/${TAG}_bench.log 2>&1
ls -la gpurun_out | tail -8
