#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/c9_bench_n2.json 2> gpurun_out/c9_bench_n2.err; echo "bench n2 rc=$?"
cut -c1-400 gpurun_out/c9_bench_n2.json
