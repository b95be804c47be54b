#!/bin/bash
# Build kernel variants with different tuning macros on the GPU box and time each.
# usage: tools/sweep.sh "<grid> <photons> <tau>" "MINB STEPS THRESH" ...
set -e
cd "$(dirname "$0")/.."
ARGS=($1); shift
for v in "$@"; do
  set -- $v
  out=/tmp/libhyp_$1_$2_$3.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -DLUCY_MIN_BLOCKS=$1 -DLUCY_STEPS_PER_ROUND=$2 -DLUCY_SERVICE_THRESHOLD=$3 \
     -o $out hyperion_b200/csrc/hyperion_b200.cu
  echo "== MIN_BLOCKS=$1 STEPS=$2 THRESH=$3"
  HYPERION_B200_LIB=$out python tools/profile_lucy.py --grid ${ARGS[0]} --photons ${ARGS[1]} --tau ${ARGS[2]} --iters 3 --n-temp ${NTEMP:-1200} | tail -2
done
