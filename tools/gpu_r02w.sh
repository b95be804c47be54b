#!/bin/bash
# Round-2 session w: emission ids reserved by the scan kernel; defaults 1024 threads, tile first, 16 service blocks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02w
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x > ${O}_tests.log 2>&1
tail -3 ${O}_tests.log
run() { echo "== TAU=${TAU:-1} $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau ${TAU:-1} --iters 4 2>&1 | grep -v "^\[wave [0-9t]" | tail -${TAILN:-1}; }
{
run X=default
TAILN=2 run HYPERION_B200_TIMING=2
TAU=5 run X=default
TAU=0.01 run X=default
run HYPERION_B200_WAVE_EMIT=5000000
run HYPERION_B200_WAVE_EMIT=10000000
} > ${O}_sweep.log 2>&1
cat ${O}_sweep.log | tail -50
