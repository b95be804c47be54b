# round-1 (session j) measurement pass: tests, bench line, ncu launch list, full capture of the dominant kernel
set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01j_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r01j_bench.json 2> gpurun_out/r01j_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-thin --no-imaging > gpurun_out/r01j_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^flight_kernel -s 3 -c 1 -f -o gpurun_out/r01j_flight python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 > gpurun_out/r01j_flight.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flight_beam -s 1 -c 1 -f -o gpurun_out/r01j_beam python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 > gpurun_out/r01j_beam.log 2>&1
