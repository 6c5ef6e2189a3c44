#!/bin/bash
# every stage map fp16 (maps16 = 2): GPU tests, bench32 trace at level 1 vs 2, default bench
mkdir -p gpurun_out
T=r39
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_tests.log 2>&1
for v in 1 2; do
  CFEN_MAPS16=$v CFEN_BENCH_TRACE=gpurun_out/${T}_trace_m$v.csv timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/${T}_bench_m$v.json 2> gpurun_out/${T}_bench_m$v.err
done
tail -n 5 gpurun_out/${T}_tests.log; for v in 1 2; do cut -c1-120 gpurun_out/${T}_bench_m$v.json; tail -n 2 gpurun_out/${T}_bench_m$v.err; done
python - <<'PY'
import json
for f in ('parity_v3_hr4_512_b1','parity_v3_hr2_512_b1','parity_v3_hr4_1024_b1','parity_v3_hr2_128_b2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, {k:(round(d[k]['max_abs_01'],5), round(d[k]['psnr'],1)) for k in ('xr','xs','xd')})
    except Exception as e: print(f, e)
PY
