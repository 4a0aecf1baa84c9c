#!/bin/bash
# Full 1-GPU validation (gpurun --timeout 1800 -- bash scripts/gpu_validate.sh): every GPU test, smoke, default bench (with the CPU
# baseline), then the ncu passes of scripts/gpu_profile_r02.sh.  FD_PARITY_FULL=1 additionally runs the cfg3 parity test over all 20
# scheduler steps (~12 min of fp32 oracle on the GPU).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/val_tests.log 2>&1; echo "gpu tests rc=$?"
tail -n 4 gpurun_out/val_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/val_smoke.log
timeout 300 python scripts/dev_callers_bench.py > gpurun_out/val_callers.log 2>&1; echo "callers rc=$?"; tail -n 4 gpurun_out/val_callers.log
timeout 200 python scripts/gpu_png_check.py > gpurun_out/val_png.log 2>&1; echo "png rc=$?"; tail -n 3 gpurun_out/val_png.log | cut -c1-600
timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; echo "bench rc=$?"
cut -c1-3500 gpurun_out/val_bench.json
timeout 300 python scripts/dev_norm_bench.py gns > gpurun_out/val_norm_bench.log 2>&1; tail -n 8 gpurun_out/val_norm_bench.log
if [ "$1" = "profile" ]; then bash scripts/gpu_profile_r02.sh; fi
