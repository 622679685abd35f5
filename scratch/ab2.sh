#!/bin/bash
# 2-GPU A/B over env settings
run() { env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
for e in "$@"; do run "$e"; done
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('1gpu', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
