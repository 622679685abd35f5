#!/bin/bash
# A/B of the co-resident variants of the persistent kernels (missm_set_coresident; MISSM_CORESIDENT pins it):
# attention kernels, the GEMM shapes of a layer and the bench step with both.  Run under gpurun.
cd "$(dirname "$0")/.."
for v in 1 0; do
  export MISSM_CORESIDENT=$v
  echo "== MISSM_CORESIDENT=$v"
  timeout 100 python scratch/attn_time.py 2>&1 | tail -1
  timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -1
  timeout -s ABRT 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${PFX:-ab}_bench_co$v.json 2> gpurun_out/${PFX:-ab}_bench_co$v.err
  echo "bench co=$v $(cut -c44-75 gpurun_out/${PFX:-ab}_bench_co$v.json) $(grep -o 'sm_mhz.\{0,10\}' gpurun_out/${PFX:-ab}_bench_co$v.json | head -1)"
done
