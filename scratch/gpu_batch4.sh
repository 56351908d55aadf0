#!/bin/bash
export PYTHONPATH=$PWD
for cfg in "512 1" "512 2" "512 4" "512 8" "1024 2" "1024 4" "1024 8" "2048 4" "2048 8" "4096 8" "4096 16"; do
  set -- $cfg
  python bench.py --problems $1 --e2e-shards $2 --no-extras --no-cpu-baseline --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('problems $1 shards $2  device %.3f ms  e2e %.3f ms  ratio %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['ms_per_step']/d['e2e']['ms_per_step']))"
done | tee gpurun_out/r2e_e2e_shards.txt
