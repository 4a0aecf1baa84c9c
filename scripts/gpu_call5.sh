#!/bin/bash
# GPU call 5: GroupNorm v4 (320-thread CTAs, deeper loads, L2 hints), BSRNet, channel-slice convs
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm or layernorm or channel_slice or conv3x3" > gpurun_out/c5_t_kern.log 2>&1; echo "kernel tests rc=$?"
timeout 600 python -m pytest tests/test_gpu_bsrnet.py -q > gpurun_out/c5_t_bsr.log 2>&1; echo "bsrnet tests rc=$?"
timeout 300 python scripts/dev_norm_bench.py > gpurun_out/c5_norm_bench.log 2>&1; echo "norm bench rc=$?"
FD_GN_L2HINT=0 timeout 300 python scripts/dev_norm_bench.py gn > gpurun_out/c5_norm_bench_nohint.log 2>&1; echo "norm bench nohint rc=$?"
cat gpurun_out/c5_norm_bench.log; echo "--- no L2 hints"; cat gpurun_out/c5_norm_bench_nohint.log
timeout 400 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/c5_bench_new.json 2> gpurun_out/c5_bench_new.err; echo "bench new rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gn_fused_kernel' -c 3 -o gpurun_out/c5_norm python scripts/dev_norm_one.py > gpurun_out/c5_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 25 gpurun_out/c5_t_kern.log | cut -c1-300
tail -n 25 gpurun_out/c5_t_bsr.log | cut -c1-300
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/c5_bench_new.json").read().strip().splitlines()[-1])
    print("value", round(j["value"], 4), "e2e", round(j["e2e"]["value"], 4), "ms", round(j["ms_per_step"], 1), j["config"]["stage_ms_last_step"],
          "gemm", round(j["roofline"]["frac"], 3), "hbm", json.dumps(j.get("roofline_hbm", {}).get("kernels")), j["clocks"])
except Exception as e:
    print("failed", e)
PY
