#!/bin/bash
# Round-2 session i: full GPU test suite, timing of the default configuration, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02i
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 3 2>&1 | tail -${TAILN:-1}; }
{
TAILN=100 run HYPERION_B200_TIMING=1
run HYPERION_B200_TILE=26,26,26
run HYPERION_B200_WAVE_TAIL=2000000
run HYPERION_B200_WAVE_TAIL=400000
} > ${O}_sweep.log 2>&1
grep -v "^\[wave" ${O}_sweep.log | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${O}_launches.csv \
   python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 2 > ${O}_ncu2.log 2>&1
tail -2 ${O}_ncu2.log
(time timeout 1500 python -m pytest tests -x -q -m gpu) > ${O}_tests.log 2>&1
tail -8 ${O}_tests.log
