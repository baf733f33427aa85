python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-configs > gpurun_out/r02x_bench_2gpu.json 2> gpurun_out/r02x_bench_2gpu.err; echo rc=$?
tail -c 1500 gpurun_out/r02x_bench_2gpu.err
head -c 300 gpurun_out/r02x_bench_2gpu.json
