#!/bin/bash
# 2-GPU bench under different NCCL settings (bucket views on); one line per variant
run() {
  tag=$1; shift
  env "$@" timeout -s ABRT 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NGPU:-2} --steps 10 --warmup 4 > gpurun_out/${PFX}_$tag.json 2> gpurun_out/${PFX}_$tag.err
  echo "$tag rc=$? $(grep -o '"value.\{0,30\}' gpurun_out/${PFX}_$tag.json | head -1) $(grep -o 'gemm_ms_per_step.\{0,12\}' gpurun_out/${PFX}_$tag.json)"
}
