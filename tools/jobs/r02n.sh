#!/bin/bash
# 8-GPU box: PCIe ceilings at 1/2/4/8 concurrent ranks, then the bench at 8 and 4 GPUs (headline only)
for n in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 tools/h2d_peak.py > gpurun_out/r02n_h2d_peak_${n}gpu.json 2> gpurun_out/r02n_h2d_${n}.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --no-configs > gpurun_out/r02n_bench_8gpu.json 2> gpurun_out/r02n_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --no-configs > gpurun_out/r02n_bench_4gpu.json 2> gpurun_out/r02n_bench4.err
nvidia-smi topo -m > gpurun_out/r02n_topo.txt 2>&1; lscpu | head -25 > gpurun_out/r02n_lscpu.txt; numactl -H > gpurun_out/r02n_numa.txt 2>&1; free -g > gpurun_out/r02n_free.txt
