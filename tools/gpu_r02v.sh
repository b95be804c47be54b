#!/bin/bash
# Round-2 session v: 896 threads with co-resident interactions against 1024 threads, tile kernel first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02v
run() { echo "== TAU=${TAU:-1} $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau ${TAU:-1} --iters 4 2>&1 | grep -v "^\[wave [0-9t]" | tail -${TAILN:-2}; }
{
for t in 1 5 0.01; do
export TAU=$t
run HYPERION_B200_WAVE_THREADS=896
run HYPERION_B200_WAVE_THREADS=1024
run HYPERION_B200_WAVE_THREADS=1024 HYPERION_B200_WAVE_SERVICE=16
done
TAU=1 run HYPERION_B200_WAVE_THREADS=1024 HYPERION_B200_POOL=28000000
TAU=1 run HYPERION_B200_WAVE_THREADS=1024 HYPERION_B200_WAVE_REFILL=10
TAU=1 run HYPERION_B200_WAVE_THREADS=1024 HYPERION_B200_WAVE_REFILL=14
} > ${O}_sweep.log 2>&1
cat ${O}_sweep.log | tail -50
