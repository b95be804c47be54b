#!/bin/bash
# Round-2 session an: ncu --set full of the first polar-grid flight launch of the c3 disk at 8 resident blocks per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02an
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flight_geo_kernel -s 0 -c 1 -o ${O}_geo_c3 \
   python tools/profile_config.py c3 1 > ${O}_ncu_geo_c3.log 2>&1
tail -1 ${O}_ncu_geo_c3.log
python tools/ncu_summary.py ${O}_geo_c3.ncu-rep 40 > ${O}_flight_geo_c3_first_ncu_full.txt 2>&1; rm -f ${O}_geo_c3.ncu-rep
head -42 ${O}_flight_geo_c3_first_ncu_full.txt
