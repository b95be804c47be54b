#!/bin/bash
# Round-2 session p: tile blocks of fewer threads (64 registers each) so that the interactions run beside them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02p
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "wave or vacuum or path_length" > ${O}_tests.log 2>&1
tail -3 ${O}_tests.log
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau ${TAU:-1} --iters 3 2>&1 | tail -${TAILN:-1}; }
{
run X=default
run HYPERION_B200_WAVE_ORDER=0
run HYPERION_B200_WAVE_THREADS=896
run HYPERION_B200_WAVE_THREADS=768
run HYPERION_B200_WAVE_THREADS=640
run HYPERION_B200_WAVE_THREADS=768 HYPERION_B200_WAVE_REFILL=8
run HYPERION_B200_WAVE_THREADS=768 HYPERION_B200_WAVE_TAIL=500000
TAU=5 run HYPERION_B200_WAVE_THREADS=1024
TAU=5 run HYPERION_B200_WAVE_THREADS=768
TAU=0.01 run HYPERION_B200_WAVE_THREADS=1024
TAU=0.01 run HYPERION_B200_WAVE_THREADS=768
} > ${O}_sweep.log 2>&1
grep -v "^\[wave" ${O}_sweep.log | tail -40
