#!/bin/bash
# build an A/B variant of libdsc.so:  build_variant.sh <name> [-DSC_LU=8 ...]   ->  gpurun_out/../variants/libdsc_<name>.so
# (bench.py / the tests pick it up through DSC_LIB=<path>)
set -e
name=$1; shift
cd "$(dirname "$0")/../.."
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -ffp-contract=off \
     -cudart static "$@" distropointclouds_b200/csrc/dsc_api.cu distropointclouds_b200/csrc/dsc_k1_build.cu \
     distropointclouds_b200/csrc/dsc_k1_f64.cu distropointclouds_b200/csrc/dsc_k23_cells.cu distropointclouds_b200/csrc/dsc_k6_labels.cu distropointclouds_b200/csrc/dsc_k7_voxel.cu -o variants/libdsc_$name.so
echo variants/libdsc_$name.so
