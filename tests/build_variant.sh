#!/bin/bash
# Build an experimental variant of the library for A/B timing on the GPU box:
#   tests/build_variant.sh <name> "<extra nvcc flags, e.g. -DTCS_X2=1>"
# compiles csrc/pe_tcs.cu with the extra flags and links it with the regular objects into
# pinn_elastodynamics_b200/libpinn_elasto_<name>.so (git-ignored; select it with PE_LIB_PATH).
set -e
name=$1; extra=$2
cd "$(dirname "$0")/../pinn_elastodynamics_b200/csrc"
make -s -j4 >/dev/null
mkdir -p /tmp/pe_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $extra \
     -c pe_tcs.cu -o /tmp/pe_variants/pe_tcs_$name.o 2> /tmp/pe_variants/pe_tcs_$name.ptxas.log || { cat /tmp/pe_variants/pe_tcs_$name.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpinn_elasto_$name.so pe_api.o pe_simt.o pe_optim.o pe_tc.o pe_tcp.o pe_lbfgs.o /tmp/pe_variants/pe_tcs_$name.o -lcudart
grep -E "registers|spill" /tmp/pe_variants/pe_tcs_$name.ptxas.log | grep -A1 -B1 "resid" | grep -E "Used|spill" | head -8
