#!/bin/bash
# Full GPU validation (gpurun --timeout 1200 -- bash scripts/gpu_validate.sh): every GPU test, smoke, callers bench, default bench, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/val_tests.log 2>&1; echo "gpu tests rc=$?"
tail -n 6 gpurun_out/val_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/val_smoke.log
timeout 300 python scripts/dev_callers_bench.py > gpurun_out/val_callers.log 2>&1; echo "callers rc=$?"; cat gpurun_out/val_callers.log | tail -n 4
timeout 600 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/val_bench.json
K='regex:gemm_tcgen05|attention_d64|gn_|layernorm|conv3x3|conv1x1|linear_smallm|tile_|cfg_euler|softmax_rows|batch_axis|upsample|nchw|nhwc|timestep|seam|scale_div|preprocess|postprocess|adain|wavelet|resample|axpby|unit_acc|canvas'
# launch list of OUR kernels only (~0.2 s of ncu overhead per profiled launch: -c 2000 covers the pre-step work and one denoise step)
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 2000 --csv --log-file gpurun_out/val_launches_1k.csv \
  python bench.py --workload 1k --steps 1 --warmup 0 --tiles-per-launch 1 --no-cpu-baseline > gpurun_out/val_ncu_1k.log 2>&1; echo "launch list rc=$?"
