#!/bin/bash
# Kernel time against blocks-per-thread (ASTC_B200_PASSES overrides launch_encode's choice).
for p in ${PASSES:-8 12 16}; do echo "== passes $p"; ASTC_B200_PASSES=$p python tools/quick_bench.py | grep -E "4096|8192|16384"; done
echo "== auto"; python tools/quick_bench.py | grep -E "4096|8192|16384"
