#!/usr/bin/env bash
# Builds libadelie_b200.so (sm_100a) in-tree and the CPU oracle.  Usage: ./build.sh [-v]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
HOSTCXX=/usr/bin/g++
[ -x "$HOSTCXX" ] || HOSTCXX=g++
EXTRA=""
if [ "${1:-}" = "-v" ]; then EXTRA="-Xptxas -v"; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda \
  -ccbin $HOSTCXX -Xcompiler -fPIC,-O3,-Wall,-Wno-unused-function,-fopenmp -shared $EXTRA \
  -o adelie_b200/libadelie_b200.so adelie_b200/csrc/capi.cu -lcudart -lgomp
make -s -C oracle
echo "built adelie_b200/libadelie_b200.so and oracle/liboracle.so"
