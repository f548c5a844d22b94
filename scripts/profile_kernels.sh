#!/usr/bin/env bash
# One `ncu --set full` capture per remaining kernel family (profiles/r1_kernels_*): dense KKT gemv, Gram kernels, device Jacobi, per-group
# sweep kernel, sparse kernels, GLM elementwise passes, Cox scans.  Small paths so that the whole script stays within a few GPU-minutes.
set -e
TAG=${1:-r1_kernels}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# dense Gaussian path, batched kernel: gemv_t / cov_small / pair_gram
N=200000 P=4000 L=8 $NCU -k regex:'gemv_t_kernel|cov_small_kernel|pair_gram_kernel' -s 6 -c 9 -o gpurun_out/${TAG}_dense -f \
    python scripts/explore_c2.py > gpurun_out/${TAG}_dense.log 2>&1 || true
# per-group sweep kernel (GLM geometry) + device Jacobi + GLM passes
N=125000 P=4000 L=6 REPS=1 $NCU -k regex:'pin_solve_kernel|jacobi_eigh_kernel|map_reduce_kernel' -s 40 -c 6 -o gpurun_out/${TAG}_glm -f \
    python scripts/explore_c3.py > gpurun_out/${TAG}_glm.log 2>&1 || true
# sparse CSC
N=500000 P=20000 M=2500 L=6 $NCU -k regex:'spmv_t_kernel|spcov_kernel|pin_solve_sparse_kernel|spaxpy_kernel' -s 70 -c 6 -o gpurun_out/${TAG}_sparse -f \
    python scripts/explore_c4.py > gpurun_out/${TAG}_sparse.log 2>&1 || true
# Cox scans
N=1000000 $NCU -k regex:'segscan_block_kernel|segscan_fix_kernel' -s 4 -c 3 -o gpurun_out/${TAG}_cox -f \
    python scripts/cox_bench.py > gpurun_out/${TAG}_cox.log 2>&1 || true
# the reports are large (the merge back is capped at 64 MiB): keep the text / csv pages, drop the .ncu-rep files
for part in dense glm sparse cox; do
    if [ -f gpurun_out/${TAG}_${part}.ncu-rep ]; then
        ncu -i gpurun_out/${TAG}_${part}.ncu-rep --page details > gpurun_out/${TAG}_${part}_details.txt 2>/dev/null || true
        ncu -i gpurun_out/${TAG}_${part}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${part}_raw.csv 2>/dev/null || true
        rm -f gpurun_out/${TAG}_${part}.ncu-rep
    fi
done
ls -la gpurun_out | grep ${TAG}
