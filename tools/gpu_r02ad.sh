#!/bin/bash
# Round-2 session ad: ncu --set full of the tile kernel after the crossing-loop clean-up (summarised on the box)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r02ad}
B="python bench.py --steps 2 --warmup 1 --no-thin --no-moderate --no-imaging --no-configs --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave_tile_kernel -s 28 -c 1 -o ${O}_wave_tile $B > ${O}_ncu_full.log 2>&1
tail -1 ${O}_ncu_full.log
python tools/ncu_summary.py ${O}_wave_tile.ncu-rep 160 > ${O}_wave_tile_ncu_full.txt 2>&1; rm -f ${O}_wave_tile.ncu-rep
head -5 ${O}_wave_tile_ncu_full.txt
