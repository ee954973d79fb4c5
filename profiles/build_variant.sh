#!/bin/bash
# build an A/B variant of the library: profiles/build_variant.sh <name> [-D flags ...]  ->  profiles/_build/<name>.so
cd "$(dirname "$0")/.."
name=$1; shift
P=paper_romualdi_2022_icra_centroidal-mpc-walking_b200
mkdir -p profiles/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr --extended-lambda -Xcompiler -fPIC -shared \
  -diag-suppress 550,177 -Xptxas=-v "$@" $P/csrc/cmpc_kernels.cu -o profiles/_build/$name.so -lcudart 2> profiles/_build/$name.ptxas.log
rc=$?
grep -A2 "cmpc_solve_team_kernelILi96ELi7ELi1E" profiles/_build/$name.ptxas.log | grep -E "registers|spill" | head -4
exit $rc
