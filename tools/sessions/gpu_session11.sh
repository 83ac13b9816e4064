#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-generate --no-hbm-kernels > gpurun_out/s11_bench_2gpu.json 2> gpurun_out/s11_bench_2gpu.err; echo "2gpu rc=$?"
head -c 700 gpurun_out/s11_bench_2gpu.json; echo; tail -3 gpurun_out/s11_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_check.py > gpurun_out/s11_ddp_check.txt 2>&1; tail -2 gpurun_out/s11_ddp_check.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --api dropin --steps 8 --warmup 3 --no-generate --no-hbm-kernels > gpurun_out/s11_bench_2gpu_dropin.json 2> gpurun_out/s11_bench_2gpu_dropin.err; echo "2gpu dropin rc=$?"
head -c 400 gpurun_out/s11_bench_2gpu_dropin.json; echo
