#!/bin/bash
# GPU call 7: ncu launch list of the 1k workload (our kernels only) + cfg5 batch workload smoke (fp8-rounded weights)
mkdir -p gpurun_out
K='regex:gemm_tcgen05|attention_d64|gn_|layernorm|conv3x3|conv1x1|linear_smallm|tile_|cfg_euler|softmax_rows|batch_axis|upsample|nchw|nhwc|timestep|seam|scale_div|preprocess|postprocess|adain|wavelet|resample|axpby|unit_acc|canvas'
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 5000 --csv --log-file gpurun_out/c7_launches_1k.csv \
  python bench.py --workload 1k --steps 1 --warmup 0 --tiles-per-launch 1 --no-cpu-baseline > gpurun_out/c7_ncu_1k.log 2>&1; echo "launch list rc=$?"
python scripts/agg_launches.py gpurun_out/c7_launches_1k.csv | head -40
timeout 400 python bench.py --workload batch32 --images 2 --fp8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c7_bench_batch.json 2> gpurun_out/c7_bench_batch.err; echo "batch bench rc=$?"
cut -c1-1200 gpurun_out/c7_bench_batch.json; tail -n 5 gpurun_out/c7_bench_batch.err
