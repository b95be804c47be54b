#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02k
nproc > ${O}_nproc.txt
(time timeout 1500 python -m pytest tests/test_gpu_configs.py -q -m gpu -s) > ${O}_tests.log 2>&1
tail -40 ${O}_tests.log
(time timeout 900 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err)
tail -c 6000 ${O}_bench.json; tail -5 ${O}_bench.err
