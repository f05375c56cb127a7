#!/bin/bash
# Build an experimental variant of the library for A/B timing on the GPU box:
#   tests/build_variant.sh <name> "<extra nvcc flags, e.g. -DTCF_EW=8>" [source.cu, default pe_tcf.cu]
# compiles the source with the extra flags and links it with the regular objects into
# pinn_elastodynamics_b200/libpinn_elasto_<name>.so (git-ignored; select it with PE_LIB_PATH).
set -e
name=$1; extra=$2; src=${3:-pe_tcf.cu}
cd "$(dirname "$0")/../pinn_elastodynamics_b200/csrc"
make -s -j4 >/dev/null
mkdir -p /tmp/pe_variants
obj=/tmp/pe_variants/${src%.cu}_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $extra \
     -c $src -o $obj 2> ${obj%.o}.ptxas.log || { cat ${obj%.o}.ptxas.log; exit 1; }
others=$(ls *.o | grep -v "^${src%.cu}.o$" | tr '\n' ' ')
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpinn_elasto_$name.so $others $obj -lcudart
grep -E "registers|spill" ${obj%.o}.ptxas.log | grep -B1 "Used 1[0-9][0-9]\|Used 2" | grep -E "Used|spill" | head -8
