#!/bin/bash
# Round-2 session ah: ncu --set full of the FIRST (4 M packets) generic flight launch of the c3 disk and of the c4 octree
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r02ah}
for c in c3 c4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flight_geo_kernel -s 0 -c 1 -o ${O}_geo_$c \
   python tools/profile_config.py $c 1 > ${O}_ncu_geo_$c.log 2>&1
tail -1 ${O}_ncu_geo_$c.log
python tools/ncu_summary.py ${O}_geo_$c.ncu-rep 60 > ${O}_flight_geo_${c}_first_ncu_full.txt 2>&1; rm -f ${O}_geo_$c.ncu-rep
head -42 ${O}_flight_geo_${c}_first_ncu_full.txt
done
