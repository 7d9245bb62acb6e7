#!/bin/bash
# build a variant of the library with extra -D flags into variants/<name>/ (development A/B tests)
name=$1; shift
mkdir -p variants/$name
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off --shared "$@" -o variants/$name/liblanczos_b200.so lanczos_hls_b200/csrc/*.cu lanczos_hls_b200/csrc/*.cpp
