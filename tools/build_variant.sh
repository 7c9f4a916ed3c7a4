#!/bin/bash
# Builds variants/lib_<tag>.so = the library with extra -D flags on kernels.cu (A/B runs: SARLACC_LIB=variants/lib_<tag>.so).
# usage: tools/build_variant.sh tag "-DSARLACC_WF_BLOCKS_LARGE=3 ..."
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p variants
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-pthread -Iinclude -Isarlacc_b200/csrc"
nvcc $F -Xptxas -v "$@" -c -o variants/kernels_$tag.o sarlacc_b200/csrc/kernels.cu 2> variants/$tag.log
nvcc -shared -o variants/lib_$tag.so variants/kernels_$tag.o sarlacc_b200/csrc/api.o sarlacc_b200/csrc/umi.o sarlacc_b200/csrc/threshold.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -pthread
rm -f variants/kernels_$tag.o
grep -A2 "wf_forward2ILi\(18\|22\)E" variants/$tag.log | grep -E "spill" | tr '\n' ' '; echo
