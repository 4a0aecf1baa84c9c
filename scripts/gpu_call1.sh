#!/bin/bash
# round-1 re-entry, GPU call 1: new norm / resample kernels — parity, microbench, A/B of the 4K bench, ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm or layernorm" > gpurun_out/c1_t_norm.log 2>&1; echo "norm tests rc=$?"
timeout 300 python -m pytest tests/test_gpu_blend_image.py -q -k "lanczos" > gpurun_out/c1_t_lz.log 2>&1; echo "lanczos tests rc=$?"
timeout 300 python scripts/dev_norm_bench.py > gpurun_out/c1_norm_bench.log 2>&1; echo "norm bench rc=$?"
FD_LN_BULK=0 timeout 120 python scripts/dev_norm_bench.py ln > gpurun_out/c1_norm_bench_ln_old.log 2>&1; echo "ln old rc=$?"
FD_GN_FUSED=0 FD_LN_BULK=0 timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c1_bench_old.json 2> gpurun_out/c1_bench_old.err; echo "bench old rc=$?"
timeout 400 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c1_bench_new.json 2> gpurun_out/c1_bench_new.err; echo "bench new rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gn_fused_kernel|layernorm_bulk_kernel' -c 5 -o gpurun_out/c1_norm python scripts/dev_norm_one.py > gpurun_out/c1_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/c1_t_norm.log gpurun_out/c1_t_lz.log
cat gpurun_out/c1_norm_bench.log gpurun_out/c1_norm_bench_ln_old.log
cat gpurun_out/c1_bench_old.json gpurun_out/c1_bench_new.json | cut -c1-1500
