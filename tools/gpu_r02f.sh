#!/bin/bash
# Round-2 session f/g: bisect of the tile-kernel regression (probe builds under build/variants)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02g
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 3 2>&1 | tail -${TAILN:-1}; }
{
run X=product
for so in build/variants/*.so; do run HYPERION_B200_LIB=$so; done
for so in build/variants/*.so; do run HYPERION_B200_LIB=$so HYPERION_B200_WAVE_QUEUE=0 HYPERION_B200_WAVE_REFILL=8 HYPERION_B200_WAVE_EMIT=25165824; done
} > ${O}_sweep.log 2>&1
grep -v "^\[wave" ${O}_sweep.log | tail -40
