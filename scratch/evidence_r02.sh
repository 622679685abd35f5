#!/bin/bash
# round-2 evidence: DRAM traffic / GB/s of the memory-bound kernels at the bench shapes (ncu, cold-cache, serialised),
# optimizer variants of the bench, the eager-torch bar at B = 64
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active"
K='regex:layernorm|colsum|compact|fusion_sum|patchify|attn_small|attn_delta|cast_f32|embed_bwd|cls_rows|scatter_rows|gather_rows|reduce_|l2norm|frame_mean|adam'
timeout 600 ncu --profile-from-start off --metrics $M -k "$K" --clock-control none --csv --log-file gpurun_out/r02i_membound_config1.csv python scratch/step_prof.py --layers 1 > gpurun_out/r02i_ncu1.log 2>&1
python scratch/membound_summary.py gpurun_out/r02i_membound_config1.csv > gpurun_out/r02i_membound_config1_summary.txt; cat gpurun_out/r02i_membound_config1_summary.txt
timeout 600 ncu --profile-from-start off --metrics $M -k "$K" --clock-control none --csv --log-file gpurun_out/r02i_membound_config2.csv python scratch/step_prof.py --layers 1 --modals audio,video --batch 32 > gpurun_out/r02i_ncu2.log 2>&1
python scratch/membound_summary.py gpurun_out/r02i_membound_config2.csv > gpurun_out/r02i_membound_config2_summary.txt; cat gpurun_out/r02i_membound_config2_summary.txt
for o in fused torch; do
  timeout -s ABRT 300 python -X faulthandler bench.py --steps 10 --warmup 3 --optimizer $o --no-cpu-baseline > gpurun_out/r02i_bench_opt_$o.json 2> gpurun_out/r02i_bench_opt_$o.err; echo "optimizer $o rc=$? $(cut -c44-75 gpurun_out/r02i_bench_opt_$o.json)"
done
timeout 600 python scratch/eager_bar.py --batch 64 --steps 3 > gpurun_out/r02i_eager_bar_b64.json 2> gpurun_out/r02i_eager.err; cat gpurun_out/r02i_eager_bar_b64.json
