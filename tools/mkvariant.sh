#!/bin/bash
# usage: mkvariant.sh NAME path/to/sls_solve.cu [extra nvcc flags]  -> scratch/variants/libNAME.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; SRC=$2; shift 2
L=systemlevelcontrol.jl_b200/lib
cp $SRC /tmp/sls_solve_var.cu
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -I systemlevelcontrol.jl_b200/csrc -c /tmp/sls_solve_var.cu -o /tmp/sls_solve_$NAME.o 2>&1 | grep -A3 "sls_solve_kernelILi3" | grep "Used\|spill" || true
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scratch/variants/lib$NAME.so $L/sls_api.o $L/sls_gather.o /tmp/sls_solve_$NAME.o $L/sls_solve_wide.o $L/sls_solve_huge.o $L/sls_solve_general.o $L/sls_scatter.o $L/sls_sim.o $L/sls_supports.o $L/sls_peak.o -cudart static
echo built scratch/variants/lib$NAME.so
