#!/bin/bash
# Build libdemcoreg_b200.so in-tree for sm_100a.  -fmad=false: the parity-critical FP64/FP32
# expressions must not be contracted into fused multiply-adds.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -O2 -shared ${DCR_NVCC_EXTRA} \
    -o ${DCR_OUT:-libdemcoreg_b200.so} dcr_front.cu dcr_select.cu dcr_bsel.cu dcr_nuth.cu dcr_corr.cu dcr_sad.cu dcr_tilt.cu dcr_comm.cu dcr_resample.cu dcr_aux.cu -ldl
