#!/bin/bash
# Round-2 session z: kernels of a round launched before the host has read the counts of its sort
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02z
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mono.py -q -m gpu -x > ${O}_tests.log 2>&1
tail -3 ${O}_tests.log
run() { echo "== TAU=${TAU:-1} $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau ${TAU:-1} --iters 4 2>&1 | grep -v "^\[wave [0-9t]" | tail -${TAILN:-1}; }
{
run X=default
run HYPERION_B200_WAVE_SPEC=0
run X=default
run HYPERION_B200_WAVE_SPEC=0
TAU=5 run X=default
TAU=5 run HYPERION_B200_WAVE_SPEC=0
TAU=0.01 run X=default
TAU=0.01 run HYPERION_B200_WAVE_SPEC=0
run HYPERION_B200_WAVE_TAIL=1500000
run HYPERION_B200_WAVE_TAIL=700000
} > ${O}_sweep.log 2>&1
cat ${O}_sweep.log | tail -50
