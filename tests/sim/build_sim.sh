#!/bin/sh
# TEST-ONLY: compile the kernels for the CPU fiber simulator (cusim.h). Not a product artefact.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC="$HERE/../../compressjs_flattened_b200/csrc"
g++ -x c++ -std=c++17 -O1 -g -fPIC -shared -DBZ_SIM -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable \
    -o "$HERE/libbz2b200_sim.so" "$SRC/bz2b200.cu"
