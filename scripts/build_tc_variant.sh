#!/bin/bash
# Build a variant of libbnuts.so with extra -D switches for logistic_tc.cu only (experiments; see the switch list at the top
# of csrc/logistic_tc.cu):   scripts/build_tc_variant.sh <tag> [-DBNUTS_TC_DEBUG=1 ...]   ->  build/libbnuts_<tag>.so
set -e
tag=$1; shift
cd "$(dirname "$0")/../inplacedhmc.jl_b200/csrc"
mkdir -p ../../build
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c -o ../../build/logistic_tc_$tag.o logistic_tc.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/libbnuts_$tag.so engine_cuda.o ../../build/logistic_tc_$tag.o gauss_tc.o
echo build/libbnuts_$tag.so
