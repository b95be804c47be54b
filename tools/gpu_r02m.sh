#!/bin/bash
# Round-2 measurement pass: bench line, launch list + DRAM traffic of the same command, ncu --set full of the
# tile kernel, full GPU test suite, smoke; probes of the box (h5py, mpirun)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r02m}
{ python -c "import h5py; print('h5py', h5py.__version__)" 2>&1 | tail -1; which h5dump mpirun mpiexec srun gfortran 2>&1; nproc; nvidia-smi -L; } > ${O}_probe.txt 2>&1
(time timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err); tail -2 ${O}_bench.err
B="python bench.py --steps 2 --warmup 1 --no-thin --no-moderate --no-imaging --no-configs --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
   --log-file ${O}_launches.csv $B > ${O}_ncu_launches.log 2>&1
tail -1 ${O}_ncu_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave_tile_kernel -s 40 -c 1 -o ${O}_wave_tile $B > ${O}_ncu_full.log 2>&1
tail -1 ${O}_ncu_full.log
(time timeout 2400 python -m pytest tests -q -m gpu) > ${O}_gpu_tests.log 2>&1
tail -6 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
