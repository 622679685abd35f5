#!/bin/bash
# usage: ddp_ab.sh NGPU "ENV1=.." "ENV2=.."   -- bench.py under torchrun for each env setting
N=$1; shift
run() { env $1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('$1', 'N=$N', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved']), d['clocks']['sm_mhz'])"; }
for e in "$@"; do run "$e"; done
