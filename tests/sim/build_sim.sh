#!/bin/sh
# TEST-ONLY: compile the kernels for the CPU fiber simulator (cusim.h). Not a product artefact.
#   SIM_SANITIZE=address|undefined tests/sim/build_sim.sh   ->  libbz2b200_sim_<sanitizer>.so next to the plain build; run with
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
#   BZ2B200_SIM_LIB=tests/sim/libbz2b200_sim_address.so python tests/sim_stress_decode.py      (libubsan.so for undefined)
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC="$HERE/../../compressjs_flattened_b200/csrc"
OUT="$HERE/libbz2b200_sim.so"
SAN=""
if [ -n "$SIM_SANITIZE" ]; then
  OUT="$HERE/libbz2b200_sim_$SIM_SANITIZE.so"
  SAN="-fsanitize=$SIM_SANITIZE -fno-omit-frame-pointer"
fi
g++ -x c++ -std=c++17 -O1 -g -fPIC -shared -DBZ_SIM $SAN -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable \
    -o "$OUT" "$SRC/bz2b200.cu"
