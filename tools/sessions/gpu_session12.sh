#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --model tv2o-large --events 4096 --batch 8 --steps 5 --warmup 3 --no-generate --no-hbm-kernels --no-cpu-baseline > gpurun_out/s12_bench_large_8gpu.json 2> gpurun_out/s12_bench_large_8gpu.err; echo "large 8gpu rc=$?"
tail -c 1200 gpurun_out/s12_bench_large_8gpu.json | head -c 700; echo; tail -3 gpurun_out/s12_bench_large_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-generate --no-hbm-kernels --no-cpu-baseline > gpurun_out/s12_bench_8gpu.json 2> gpurun_out/s12_bench_8gpu.err; echo "medium 8gpu rc=$?"
head -c 500 gpurun_out/s12_bench_8gpu.json; echo
