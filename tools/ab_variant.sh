#!/bin/bash
# Build the current csrc/ as an experiment library: tools/ab_variant.sh NAME  -> ofdm-course_b200/lib/exp/NAME.so
# (only chain_rx4096.cu and api.cu are recompiled; the other objects are shared with the main build)
set -e
cd "$(dirname "$0")/../ofdm-course_b200"
make -s >/dev/null
mkdir -p lib/exp
cp lib/libofdm_b200.so lib/exp/$1.so
grep -A2 "rx4096_kernelILb1ELb0" build/chain_rx4096.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
