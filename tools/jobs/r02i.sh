#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02i_tests.log
python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
