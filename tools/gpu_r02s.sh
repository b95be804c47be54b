#!/bin/bash
# Round-2 session s: keys by position (coalesced sort), 896-thread tile blocks with co-resident interactions
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02s
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > ${O}_tests.log 2>&1
tail -3 ${O}_tests.log
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau ${TAU:-1} --iters 3 2>&1 | grep -v "^\[wave [0-9t]" | tail -${TAILN:-1}; }
{
run X=default
TAILN=2 run HYPERION_B200_TIMING=2
run HYPERION_B200_WAVE_THREADS=1024 HYPERION_B200_WAVE_ORDER=0
run HYPERION_B200_WAVE_THREADS=1024 HYPERION_B200_WAVE_ORDER=1
run HYPERION_B200_WAVE_TAIL=500000
run HYPERION_B200_WAVE_TAIL=2000000
TAU=5 run X=default
TAU=0.01 run X=default
} > ${O}_sweep.log 2>&1
cat ${O}_sweep.log | tail -40
