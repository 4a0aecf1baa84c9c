#!/bin/bash
# GPU call 6: full validation — every GPU test, smoke, callers bench, default bench, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c6_tests.log 2>&1; echo "gpu tests rc=$?"
tail -n 6 gpurun_out/c6_tests.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/c6_smoke.log
timeout 300 python scripts/dev_callers_bench.py > gpurun_out/c6_callers.log 2>&1; echo "callers rc=$?"; cat gpurun_out/c6_callers.log | tail -n 4
timeout 600 python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/c6_bench.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/c6_launches_1k.csv \
  python bench.py --workload 1k --steps 1 --warmup 0 --tiles-per-launch 1 --no-cpu-baseline > gpurun_out/c6_ncu_1k.log 2>&1; echo "launch list rc=$?"
