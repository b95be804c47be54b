#!/bin/bash
# Round-2 closing pass with the final library: bench line, full GPU test suite, smoke
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02zzb
(time timeout 1200 python bench.py --steps 5 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err); tail -2 ${O}_bench.err
(time timeout 2400 python -m pytest tests -q -m gpu) > ${O}_gpu_tests.log 2>&1
tail -6 ${O}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err; tail -c 600 ${O}_bench_ref.json
