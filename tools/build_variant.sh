#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh <name> [-D... nvcc flags]  -> build/lib_<name>.so
# (run a tool or bench.py against it with EAQHM_LIB=build/lib_<name>.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
     -o build/lib_$name.so eaqhm-analysis-and-synthesis-in-python_b200/csrc/*.cu
echo build/lib_$name.so
