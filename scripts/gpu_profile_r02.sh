#!/bin/bash
# Round-2 profiling pass (gpurun --timeout 1500 -- bash scripts/gpu_profile_r02.sh).  Each ncu step runs only after the same
# command exited 0 without ncu; numbers printed under ncu are never bench values.
mkdir -p gpurun_out
# 1. top GEMM / conv shapes at the bench's launch size (14 samples): CUDA-event TFLOP/s, then ncu --set full of the same launches
timeout 200 python scripts/dev_gemm_top_shapes.py 14 > gpurun_out/r02_gemm_top_shapes.log 2>&1; echo "top shapes rc=$?"; cat gpurun_out/r02_gemm_top_shapes.log | tail -n 9
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -o gpurun_out/r02_gemm_top_shapes \
  python scripts/dev_gemm_top_shapes.py 14 > gpurun_out/r02_gemm_top_shapes_ncu.log 2>&1; echo "ncu top shapes rc=$?"
ncu -i gpurun_out/r02_gemm_top_shapes.ncu-rep --page raw --csv > gpurun_out/r02_ncu_gemm_top_shapes_raw.csv 2>/dev/null
# 2. unfiltered launch list (time + DRAM bytes of EVERY kernel) of one 7-tile eager UNet evaluation + the fused step tail
timeout 300 python scripts/dev_tilestep_eager.py 7 > gpurun_out/r02_tilestep_eager.log 2>&1; echo "tilestep eager rc=$?"; tail -n 1 gpurun_out/r02_tilestep_eager.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_tilestep7_unfiltered.csv python scripts/dev_tilestep_eager.py 7 > gpurun_out/r02_tilestep_eager_ncu.log 2>&1; echo "ncu unfiltered rc=$?"
python scripts/agg_traffic.py gpurun_out/r02_launches_tilestep7_unfiltered.csv gpurun_out/r02_gemm_alg.json gpurun_out/r02_gemm_traffic.json > gpurun_out/r02_launches_tilestep7_unfiltered_summary.txt 2>&1
head -n 30 gpurun_out/r02_launches_tilestep7_unfiltered_summary.txt
# 3. unfiltered share of every kernel in one whole 20-step 4K restoration (CUPTI activity records through torch.profiler)
timeout 400 python scripts/dev_profile_shares.py 4k 20 > gpurun_out/r02_kernel_shares_4k.txt 2>&1; echo "shares rc=$?"; head -n 25 gpurun_out/r02_kernel_shares_4k.txt
# 4. long-KV attention at the 4K bench's shape (14 samples x 10 heads x 4096 keys): --set full with source-level stall sampling
timeout 120 python scripts/dev_attn_one.py 14 10 4096 4096 3 > gpurun_out/r02_attn_one.log 2>&1; echo "attn one rc=$?"; cat gpurun_out/r02_attn_one.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:attention_d64_kernel --launch-skip 3 -c 1 -o gpurun_out/r02_attn4096 \
  python scripts/dev_attn_one.py 14 10 4096 4096 3 > gpurun_out/r02_attn_ncu.log 2>&1; echo "ncu attention rc=$?"
ncu -i gpurun_out/r02_attn4096.ncu-rep --page raw --csv > gpurun_out/r02_ncu_attention_4096_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_attn4096.ncu-rep --page source --csv > gpurun_out/r02_ncu_attention_4096_source.csv 2>/dev/null
# 5. d=512 attention (VAE / LQ-encoder mid block): TFLOP/s next to the round-1 GEMM-softmax-GEMM path, then one --set full capture
timeout 200 python scripts/dev_attn512.py > gpurun_out/r02_attn512.log 2>&1; echo "attn512 rc=$?"; tail -n 6 gpurun_out/r02_attn512.log
timeout 400 ncu --set full --import-source on --clock-control none -k regex:attention_d512_kernel --launch-skip 4 -c 1 -o gpurun_out/r02_attn512 \
  python scripts/dev_attn512.py > gpurun_out/r02_attn512_ncu.log 2>&1; echo "ncu attn512 rc=$?"
ncu -i gpurun_out/r02_attn512.ncu-rep --page raw --csv > gpurun_out/r02_ncu_attention_d512_raw.csv 2>/dev/null
# 6. GroupNorm from producer statistics (finalize + streaming apply) at the tile-step's largest shape, L2 flushed
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gn_apply_stream_kernel --launch-skip 6 -c 1 -o gpurun_out/r02_gn_apply \
  python scripts/dev_norm_bench.py gns > gpurun_out/r02_gn_apply_ncu.log 2>&1; echo "ncu gn apply rc=$?"
ncu -i gpurun_out/r02_gn_apply.ncu-rep --page raw --csv > gpurun_out/r02_ncu_gn_apply_raw.csv 2>/dev/null
