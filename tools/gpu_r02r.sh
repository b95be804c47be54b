#!/bin/bash
# ncu --set full of the service kernels of the wave engine
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02r
P="python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 2"
export HYPERION_B200_WAVE_ORDER=0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_interact_kernel -s 26 -c 1 -o ${O}_interact $P > ${O}_ncu1.log 2>&1; tail -1 ${O}_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_scatter_kernel -s 30 -c 1 -o ${O}_scatter $P > ${O}_ncu2.log 2>&1; tail -1 ${O}_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_emit_kernel -s 1 -c 1 -o ${O}_emit $P > ${O}_ncu3.log 2>&1; tail -1 ${O}_ncu3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wave_tile_kernel -s 28 -c 1 -o ${O}_tile $P > ${O}_ncu4.log 2>&1; tail -1 ${O}_ncu4.log
for k in interact scatter emit tile; do python tools/ncu_summary.py ${O}_$k.ncu-rep 40 > ${O}_${k}_ncu_full.txt 2>&1; rm -f ${O}_$k.ncu-rep; done
ls -la gpurun_out
