# round-1 (session j) measurement pass: tests, bench line, ncu launch list (+ DRAM bytes), full captures of the
# two flight kernels summarised on the box (the .ncu-rep files are too large to bring back)
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r01l_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r01l_bench.json 2> gpurun_out/r01l_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01l_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-thin --no-imaging > gpurun_out/r01l_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^flight_kernel -s 3 -c 1 -f -o /tmp/r01l_flight python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 > gpurun_out/r01l_flight.log 2>&1
python tools/ncu_summary.py /tmp/r01l_flight.ncu-rep 40 > gpurun_out/r01l_flight_kernel_ncu_full.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flight_beam -s 1 -c 1 -f -o /tmp/r01l_beam python tools/profile_lucy.py --grid 256 --photons 2e7 --tau 1 --iters 1 > gpurun_out/r01l_beam.log 2>&1
python tools/ncu_summary.py /tmp/r01l_beam.ncu-rep 40 > gpurun_out/r01l_flight_beam_ncu_full.txt 2>&1
ls -la gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r01l_smoke.log 2>&1
