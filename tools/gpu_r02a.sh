#!/bin/bash
# Round-2 session a: first run of the wave engine on the GPU (parity, then timing sweep, then one ncu capture).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > ${O}_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "wave or path_length or lucy" > ${O}_tests.log 2>&1
tail -5 ${O}_tests.log
run() { echo "== $*"; env "$@" timeout 300 python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 3 2>&1 | tail -${TAILN:-2}; }
{
run HYPERION_B200_ENGINE=rounds
TAILN=60 run HYPERION_B200_TIMING=1
run HYPERION_B200_WAVE_CHUNK=4096
run HYPERION_B200_WAVE_CHUNK=8192
run HYPERION_B200_POOL=8388608
run HYPERION_B200_POOL=16777216
run HYPERION_B200_POOL=25165824
run HYPERION_B200_TILE=32,26,26
run HYPERION_B200_TILE=22,22,22
run HYPERION_B200_TILE=16,16,16
run HYPERION_B200_WAVE_TAIL=1000000
run HYPERION_B200_WAVE_TAIL=100000
} > ${O}_sweep.log 2>&1
cat ${O}_sweep.log | grep -v "^\[wave" | tail -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_tile_kernel -s 6 -c 1 -o ${O}_wave_tile \
   python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
