python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/bench_init_preint.py 2>&1 | tail -1 | tee gpurun_out/r02y_init_preint.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --no-configs --steps 5 --warmup 3 > gpurun_out/r02y_bench_2gpu.json 2> gpurun_out/r02y_bench_2gpu.err; echo rc=$?
head -c 120 gpurun_out/r02y_bench_2gpu.json; echo; wc -l gpurun_out/r02y_bench_2gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | head -c 300
