#!/bin/bash
# First GPU call of the next round (gpurun --timeout 900 -- bash scripts/gpu_experimental.sh): acceptance tests and A/B numbers of
# the opt-in paths written at the end of round 1 — LayerNorm folding (FD_LN_FOLD=1), 8-softmax-warp attention (FD_ATTN_W8=1).
# The peer-memory exchange (FD_P2P_EXCHANGE=1) needs 2 GPUs:
#   gpurun --gpus 2 --timeout 400 -- 'FD_TEST_EXPERIMENTAL=1 timeout 350 python -m pytest tests/test_gpu_multi.py -q'
# Every step runs under its own `timeout`: a protocol error in a new kernel is a hang, not a wrong number.
mkdir -p gpurun_out
FD_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_ln_fold.py tests/test_gpu_attention_variants.py -q > gpurun_out/exp_tests.log 2>&1; echo "experimental tests rc=$?"
tail -n 30 gpurun_out/exp_tests.log | cut -c1-300
for v in "" "FD_ATTN_POLY=3" "FD_ATTN_W8=1" "FD_ATTN_W8=1 FD_ATTN_POLY=3"; do
  env $v timeout 60 python scripts/dev_attn_poly.py 2>&1 | grep TFLOP
done
# UNet parity with the fold on (full-size network vs the fp32 oracle) and the 4K bench A/B
FD_LN_FOLD=1 timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_unet_pipeline.py -q -x > gpurun_out/exp_fold_parity.log 2>&1; echo "parity with FD_LN_FOLD=1 rc=$?"
tail -n 5 gpurun_out/exp_fold_parity.log | cut -c1-300
timeout 400 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/exp_bench_default.json 2> gpurun_out/exp_bench_default.err; echo "bench default rc=$?"
FD_LN_FOLD=1 timeout 400 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/exp_bench_fold.json 2> gpurun_out/exp_bench_fold.err; echo "bench fold rc=$?"
FD_ATTN_W8=1 timeout 400 python bench.py --steps 1 --warmup 2 --no-cpu-baseline > gpurun_out/exp_bench_w8.json 2> gpurun_out/exp_bench_w8.err; echo "bench w8 rc=$?"
python - <<'PY'
import json
for n in ("default", "fold", "w8"):
    try:
        j = json.loads(open("gpurun_out/exp_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value", round(j["value"], 4), "ms", round(j["ms_per_step"], 1), j["config"]["stage_ms_last_step"], "tile-step ms", round(j["roofline"]["unet_tile_step_ms"], 3))
    except Exception as e:
        print(n, "failed", e)
PY
