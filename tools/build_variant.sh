#!/bin/bash
# Development builds of the library with extra -D flags:  tools/build_variant.sh <name> [-DFLAG ...]  ->  build/libdeepsir_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
src=deepsir_b200/csrc
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  $src/api.cu $src/knn.cu $src/knn_grid.cu $src/knn_tree.cu $src/match_fp32.cu $src/match_tc.cu $src/match_tc_soft.cu \
  $src/kabsch.cu $src/graph.cu $src/keypoint.cu $src/metrics.cu -o build/libdeepsir_$name.so 2>&1 | grep -E "error" || true
ls -la build/libdeepsir_$name.so
