#!/bin/bash
# A/B builds of the CUDA library: build/libvp9b200_<name>.so compiled with extra -D flags
# (select with VP9B_LIB=build/libvp9b200_<name>.so).   usage: build_variant.sh <name> [-DVP9B_...=...]...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build
C=cuda-vp9_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-sign-compare,-march=x86-64-v3 -shared -Iinclude -I$C "$@" \
  $C/vp9_recon.cu $C/vp9_parser.cc $C/capi_host.cc $C/vp9_decoder.cc $C/vp9_packer.cc -o build/libvp9b200_$name.so
cuobjdump -res-usage build/libvp9b200_$name.so | grep -A1 "k_lf_sbIhLb1ENS_11FrameBatchTILi32\|k_inter_pred_sbIhNS_11FrameBatchTILi32\|k_intraIhiNS_11FrameBatchTILi32EEELi4" | grep -o "k_[a-z_]*\|REG:[0-9]*\|STACK:[0-9]*" | paste - - - 
