#!/bin/bash
# final-build evidence: GPU tests, bench, reference arm, ncu launch list of the bench command, ncu --set full of every kernel
python -m pytest tests -m gpu -q > gpurun_out/r02x_gpu_tests.log 2>&1; tail -3 gpurun_out/r02x_gpu_tests.log
python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err || tail -5 gpurun_out/r02x_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02x_bench_reference_arm.json 2>> gpurun_out/r02x_bench.err
python tools/profile_driver.py > /dev/null 2>&1 && echo driver-ok
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu > gpurun_out/r02x_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:"marg_|preintegrate|zero_" -c 40 -f -o /tmp/r02x_prof python tools/profile_driver.py > gpurun_out/r02x_ncu_full.log 2>&1
tail -2 gpurun_out/r02x_ncu_full.log
python tools/ncu_summary.py /tmp/r02x_prof.ncu-rep gpurun_out/r02x_ncu_full_summary.json > /dev/null 2>&1 && echo summary-ok
python tools/sass_census.py > gpurun_out/r02x_sass_census.json 2>/dev/null && echo sass-ok
ls -la /tmp/r02x_prof.ncu-rep; du -sh gpurun_out
