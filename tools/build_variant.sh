#!/bin/sh
# tools/build_variant.sh NAME "-DFLAG=..." -> oblas_b200/liboblas_b200_NAME.so (a same-box A/B build: OB2_LIB=<that file> python tools/tc_classes.py)
set -e
cd "$(dirname "$0")/../oblas_b200/csrc"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
  --expt-relaxed-constexpr -Xptxas -v $2 -shared -o ../liboblas_b200_$1.so oblas_b200.cu tables_export.cc 2> /tmp/ptxas_$1.log || { cat /tmp/ptxas_$1.log; exit 1; }
grep -A3 "update_tc_kernelILb0" /tmp/ptxas_$1.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $1"
