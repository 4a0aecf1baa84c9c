#!/bin/bash
# GPU call 4: packed-math GroupNorm (fast SiLU) and LayerNorm
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -q -k "groupnorm or layernorm" > gpurun_out/c4_t_norm.log 2>&1; echo "norm tests rc=$?"
timeout 300 python scripts/dev_norm_bench.py > gpurun_out/c4_norm_bench.log 2>&1; echo "norm bench rc=$?"
cat gpurun_out/c4_norm_bench.log
timeout 400 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/c4_bench_new.json 2> gpurun_out/c4_bench_new.err; echo "bench new rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gn_fused_kernel|layernorm_kernel' -c 5 -o gpurun_out/c4_norm python scripts/dev_norm_one.py > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/c4_t_norm.log
python - <<'PY'
import json
for n in ("new",):
    try:
        j = json.loads(open("gpurun_out/c4_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value", round(j["value"], 4), "e2e", round(j["e2e"]["value"], 4), "ms", round(j["ms_per_step"], 1), j["config"]["stage_ms_last_step"],
              "gemm", round(j["roofline"]["frac"], 3), "hbm", json.dumps(j.get("roofline_hbm", {}).get("kernels")), j["clocks"])
    except Exception as e:
        print(n, "failed", e)
PY
