#!/bin/bash
# 2-GPU session of the final code: multi-GPU boundary tests (incl. monochromatic + PDA), bench at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02zy
nvidia-smi -L > ${O}_gpus.txt
(time timeout 1200 python -m pytest tests/test_gpu_multigpu.py -q -m gpu -s) > ${O}_tests.log 2>&1
tail -15 ${O}_tests.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 5 --warmup 3 > ${O}_bench_n2.json 2> ${O}_bench_n2.err); tail -3 ${O}_bench_n2.err
python - <<'PY'
import json
b = json.load(open("gpurun_out/r02zy_bench_n2.json"))
print({k: b.get(k) for k in ("value", "ms_per_step", "n_gpus", "multi_gpu_check")}, b["e2e"], b["roofline"]["frac"])
PY
