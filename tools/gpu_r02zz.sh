#!/bin/bash
# Round-2 final measurement pass: bench line, launch list + DRAM traffic of the same command, ncu --set full of the
# tile kernel and of the generic flight kernel on the c3 disk (summarised on the box), full GPU test suite, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r02zz}
(time timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err); tail -2 ${O}_bench.err
B="python bench.py --steps 2 --warmup 1 --no-thin --no-moderate --no-imaging --no-configs --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2600 --csv \
   --log-file ${O}_launches.csv $B > ${O}_ncu_launches.log 2>&1
tail -1 ${O}_ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave_tile_kernel -s 28 -c 1 -o ${O}_wave_tile $B > ${O}_ncu_full.log 2>&1
tail -1 ${O}_ncu_full.log
python tools/ncu_summary.py ${O}_wave_tile.ncu-rep 40 > ${O}_wave_tile_ncu_full.txt 2>&1; rm -f ${O}_wave_tile.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flight_geo_kernel -s 3 -c 1 -o ${O}_geo_c3 \
   python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-thin --no-moderate --no-imaging --no-configs > ${O}_ncu_geo.log 2>&1
tail -1 ${O}_ncu_geo.log
python tools/ncu_summary.py ${O}_geo_c3.ncu-rep 40 > ${O}_flight_geo_c3_ncu_full.txt 2>&1; rm -f ${O}_geo_c3.ncu-rep
(time timeout 2400 python -m pytest tests -q -m gpu) > ${O}_gpu_tests.log 2>&1
tail -6 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
ls -la gpurun_out | tail -20
