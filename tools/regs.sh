#!/bin/bash
# Register / spill report of the encode kernels for a set of -D flags: tools/regs.sh -DASTC_CPASYNC_4X4=1 ...
cd "$(dirname "$0")/.."
nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr \
  -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -I include -I astc_encoder_b200/csrc \
  -Xptxas -v "$@" -c astc_encoder_b200/csrc/astc_kernels.cu -o /tmp/regs_$$.o 2>&1 \
  | awk '/Compiling entry function/ {name=$0} /registers/ {r=$0} /spill/ {sp=$0} /registers/ {print name; print "   " sp; print "   " r}' \
  | c++filt | grep -A2 -E "encode(4x4|6x6)" | grep -v "^--"
rm -f /tmp/regs_$$.o
