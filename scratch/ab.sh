#!/bin/bash
# usage: ab.sh "ENV1=.." "ENV2=.." ...   runs bench for each env setting, interleaved twice
run() { env $1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), round(d['ms_per_step'],2), round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],1), d['clocks']['sm_mhz'])"; }
for rep in 1 2; do for e in "$@"; do run "$e"; done; done
