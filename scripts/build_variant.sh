#!/bin/bash
# A/B builds of libsalsa_b200.so: scripts/build_variant.sh <tag> [-DNAME=VALUE ...] -> salsa_b200/_build/libsalsa_<tag>.so
# (select it at run time with SALSA_B200_LIB=salsa_b200/_build/libsalsa_<tag>.so; the CRNN objects are the default build's)
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p salsa_b200/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Isalsa_b200/csrc "$@" \
     -c salsa_b200/csrc/salsa_abi.cu -o salsa_b200/_build/salsa_abi_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o salsa_b200/_build/libsalsa_$tag.so salsa_b200/_build/salsa_abi_$tag.o salsa_b200/_build/crnn_abi.cu.o salsa_b200/_build/crnn_model.cu.o -lcudart
echo salsa_b200/_build/libsalsa_$tag.so
