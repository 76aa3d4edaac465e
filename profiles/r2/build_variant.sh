#!/bin/bash
# profiles/r2/build_variant.sh NAME [nvcc -D flags...]: the library with extra flags -> profiles/r2/variants/NAME.so
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
src=deep_prob_feature_track_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -I include -I $src \
  -o profiles/r2/variants/$name.so $src/*.cu
