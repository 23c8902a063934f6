#!/bin/bash
# Build a variant of libbnuts.so with extra -D switches for logistic_rm.cu only (timing experiments, see BNUTS_RM_DEBUG):
#   scripts/build_rm_variant.sh <tag> [-DBNUTS_RM_DEBUG=1 ...]   ->  build/libbnuts_<tag>.so
set -e
tag=$1; shift
cd "$(dirname "$0")/../inplacedhmc.jl_b200/csrc"
mkdir -p ../../build
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c -o ../../build/logistic_rm_$tag.o logistic_rm.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libbnuts_$tag.so engine_cuda.o logistic_tc.o ../../build/logistic_rm_$tag.o gauss_tc.o
echo build/libbnuts_$tag.so
