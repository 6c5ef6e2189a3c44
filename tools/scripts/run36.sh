#!/bin/bash
# batch-1 latency vs fusion threshold
mkdir -p gpurun_out
T=r36
for mt in 96 32 8 1; do
  CFEN_FUSE_MIN_TILES=$mt timeout 300 python bench.py --batch 32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_lat_mt$mt.json 2> gpurun_out/${T}_lat_mt$mt.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/${T}_lat_mt$mt.json')); print('min_tiles', $mt, d['latency_batch1'], d['value'])"
done
