#!/bin/bash
# Build flight-kernel variants (tuning macros) and time each on the GPU box.
# usage (here, no GPU):   tools/sweep.sh build "MINB D GROUPS" ...
#       (on the GPU box):  tools/sweep.sh run "<grid> <photons> <tau>" [env...]
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
mkdir -p build/variants
if [ "$mode" = build ]; then
  for v in "$@"; do
    set -- $v
    out=build/variants/libhyp_$1_$2_$3.so
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
       -DFLIGHT_MIN_BLOCKS=$1 -DFLIGHT_LOOKAHEAD=$2 -DFLIGHT_GROUPS=$3 \
       -o $out hyperion_b200/csrc/hyperion_b200.cu &
  done
  wait
  ls -la build/variants
else
  ARGS=($1); shift
  for so in build/variants/*.so; do
    echo "== $so $*"
    env "$@" HYPERION_B200_LIB=$so python tools/profile_lucy.py --grid ${ARGS[0]} --photons ${ARGS[1]} --tau ${ARGS[2]} --iters 3 --n-temp ${NTEMP:-1200} | tail -2
  done
fi
