#!/bin/bash
# registers / spills of the encode kernels: tools/regs.sh [-DASTC_DEV_FEW_VARIANTS ...]
cd "$(dirname "$0")/.."
nvcc -ccbin /usr/bin/g++ -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -fmad=false \
  -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -I include -I astc_encoder_b200/csrc -Xptxas -v "$@" \
  -c astc_encoder_b200/csrc/astc_kernels.cu -o /tmp/astc_kernels_regs.o 2>&1 | \
  awk '/Compiling entry function/ {name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)} /spill stores/ {sp=$0} /Used [0-9]+ registers/ {if (name ~ /encode/) print name, "|", $0, "|", sp}' | \
  sed -e 's/_ZN4astc16//' -e 's/EEvNS_12EncodeParamsE//' -e 's/ptxas info *: *//g'
