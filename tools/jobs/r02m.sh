#!/bin/bash
python tools/latency_breakdown.py 1000 1 > gpurun_out/r02m_latency.jsonl 2>gpurun_out/r02m.err
python tools/latency_breakdown.py 150 1 >> gpurun_out/r02m_latency.jsonl 2>>gpurun_out/r02m.err
python tools/latency_breakdown.py 150 512 >> gpurun_out/r02m_latency.jsonl 2>>gpurun_out/r02m.err
