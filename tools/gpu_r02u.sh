#!/bin/bash
# Round-2 session u: per-lane refill in the generic (spherical / octree / AMR) flight kernel
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/r02u
(time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_imaging.py tests/test_gpu_mrw.py tests/test_gpu_pda.py tests/test_gpu_sources.py tests/test_gpu_mono.py -q -m gpu -x) > ${O}_tests.log 2>&1
tail -5 ${O}_tests.log
for w in c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-thin --no-moderate --no-imaging --no-configs > ${O}_bench_$w.json 2> ${O}_bench_$w.err
  python - <<PY
import json
d=json.load(open("${O}_bench_$w.json"))
print("$w", d.get("value"), d.get("ms_per_step"), d.get("roofline",{}).get("frac"), {k:v for k,v in d.items() if k.startswith("imaging")})
PY
done
