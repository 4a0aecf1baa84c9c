#!/bin/bash
# GPU call 8 (2 GPUs): the one-launch GroupNorm next to NCCL — 2-rank bench (short) must not hang
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c8_bench_n2.json 2> gpurun_out/c8_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-1500 gpurun_out/c8_bench_n2.json; tail -n 4 gpurun_out/c8_bench_n2.err | cut -c1-300
