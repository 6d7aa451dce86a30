#!/bin/bash
# Instruction mix of one kernel in the built library: tools/sass_mix.sh <mangled-name-substring>
LIB=${2:-astc_encoder_b200/libastc_b200.so}
cuobjdump -sass "$LIB" | awk -v pat="$1" '/Function :/ {f = index($0, pat) > 0} f' | grep -E '^\s+/\*[0-9a-f]{4}\*/' \
  | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)(\.[A-Za-z0-9_.]+)?\s.*/\2/' | sort | uniq -c | sort -rn
