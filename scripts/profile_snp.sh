#!/usr/bin/env bash
# ncu evidence for the SNP unphased kernels (profiles/r1_snp_*): launch list of a reduced config-5 path and --set full captures of the
# packed-bit transposed GEMV (K = 8 multi-response and K = 1) and of the decode kernel.
set -e
TAG=${1:-r1_snp}
mkdir -p gpurun_out
export N=${N:-500000} P=${P:-20000} L=${L:-8} REPS=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python scripts/explore_c5.py > gpurun_out/${TAG}_launches.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:snp_gemv_t_kernel -s 1 -c 3 -o gpurun_out/${TAG}_gemv -f \
    python scripts/explore_c5.py > gpurun_out/${TAG}_gemv.log 2>&1 || true
ncu --set full --clock-control none --import-source on -k regex:snp_decode_kernel -s 70 -c 1 -o gpurun_out/${TAG}_decode -f \
    python scripts/explore_c5.py > gpurun_out/${TAG}_decode.log 2>&1 || true
tail -4 gpurun_out/${TAG}_launches.log
ls -la gpurun_out | grep ${TAG}
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_cox_launches.csv \
    python scripts/cox_bench.py > gpurun_out/${TAG}_cox.log 2>&1 || true
python scripts/cox_bench.py > gpurun_out/${TAG}_cox_unprofiled.log 2>&1 || true
