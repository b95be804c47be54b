#!/bin/bash
# Round-2 session h: wave engine after the bisect (ld.cs records, plain sort atomics, no emission throttle)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02h
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wave or path_length or lucy" > ${O}_tests.log 2>&1
tail -5 ${O}_tests.log
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 3 2>&1 | tail -${TAILN:-1}; }
{
TAILN=100 run HYPERION_B200_TIMING=1
run HYPERION_B200_WAVE_QUEUE=0
run HYPERION_B200_WAVE_REFILL=8
run HYPERION_B200_WAVE_REFILL=12
run HYPERION_B200_WAVE_REFILL=20
run HYPERION_B200_WAVE_QUEUE=0 HYPERION_B200_WAVE_REFILL=8
run HYPERION_B200_WAVE_THREADS=896
run HYPERION_B200_WAVE_THREADS=768
run HYPERION_B200_TILE=28,28,28
run HYPERION_B200_WAVE_TAIL=1000000
run HYPERION_B200_LIB=build/variants/libhyp_align128.so
run HYPERION_B200_LIB=build/variants/libhyp_align128.so HYPERION_B200_WAVE_QUEUE=0
} > ${O}_sweep.log 2>&1
grep -v "^\[wave" ${O}_sweep.log | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_tile_kernel -s 5 -c 1 -o ${O}_wave_tile \
   python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
