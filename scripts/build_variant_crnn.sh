#!/bin/bash
# A/B builds of the CRNN half of libsalsa_b200.so: scripts/build_variant_crnn.sh <tag> [-DNAME=VALUE ...]
# -> salsa_b200/_build/libsalsa_<tag>.so (select it with SALSA_B200_LIB=...; the feature objects are the default build's)
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p salsa_b200/_build
for f in crnn_abi crnn_model; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Isalsa_b200/csrc "$@" \
       -c salsa_b200/csrc/$f.cu -o salsa_b200/_build/${f}_$tag.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o salsa_b200/_build/libsalsa_$tag.so salsa_b200/_build/salsa_abi.cu.o salsa_b200/_build/crnn_abi_$tag.o salsa_b200/_build/crnn_model_$tag.o -lcudart
echo salsa_b200/_build/libsalsa_$tag.so
