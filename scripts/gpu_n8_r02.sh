#!/bin/bash
# 8-GPU measurement pass (gpurun --gpus 8 --timeout 900 -- bash scripts/gpu_n8_r02.sh): BASELINE cfg3 / cfg4 / cfg5 on one 8xB200 node
mkdir -p gpurun_out
run() {  # name, extra env, args...
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 8 "$@" > gpurun_out/r02_bench_${name}.json 2> gpurun_out/r02_bench_${name}.err; echo "$name rc=$?"
  python - "$name" <<'PY'
import json, sys
try:
    j = json.loads(open("gpurun_out/r02_bench_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(j["value"], 4), "ms", round(j["ms_per_step"], 1), "e2e", round(j["e2e"]["value"], 4), j["config"]["stage_ms_last_step"],
          "parity", j.get("parity_n_rank_vs_1_rank"), "clocks", j.get("clocks", {}).get("sm_mhz"), j.get("clocks", {}).get("reasons"))
except Exception as e:
    print(sys.argv[1], "no JSON line:", e)
PY
}
run cfg3_4k_n8 FD_X=0 --workload 4k --steps 2 --warmup 2 --no-cpu-baseline
run cfg4_8k_n8 FD_X=0 --workload 8k --steps 1 --warmup 1 --no-cpu-baseline --no-parity
run cfg5_batch32_bf16_n8 FD_X=0 --workload batch32 --steps 1 --warmup 1 --no-cpu-baseline
run cfg5_batch32_fp8_n8 FD_X=0 --workload batch32 --fp8 --steps 1 --warmup 1 --no-cpu-baseline
run cfg3_4k_n8_p2p FD_P2P_EXCHANGE=1 --workload 4k --steps 2 --warmup 1 --no-cpu-baseline --no-parity
