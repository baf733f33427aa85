#!/bin/bash
python -m pytest tests/test_forensic_gpu.py -m gpu -x -q -s 2>&1 | grep -v "^E  \|^    " | head -60 > gpurun_out/r02j_tests.log
