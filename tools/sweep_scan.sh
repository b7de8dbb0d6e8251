#!/bin/bash
# scan-kernel tuning sweep (diagnostic)
export B200VS_SCAN=ldg
summ() { python -c "
import json,sys
d=json.load(open('$1'))
r=d['roofline']
print('$2', 'qps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'kernel ms', round(r['avg_launch_ms'],4), 'GB/s', round(r['achieved']), 'frac', round(r['frac'],3))
"; }
for W in 10Mx128; do
for warps in 8 12 16; do for R in 4 8 16; do for mw in 32 48; do
  B200VS_SCAN_WARPS=$warps B200VS_SCAN_R=$R B200VS_SCAN_MAXWARPS=$mw timeout 200 python bench.py --workload $W --batch 1 --steps 30 --warmup 3 --extras 0 --cpu-baseline 0 > gpurun_out/sw.json 2>gpurun_out/sw.err && summ gpurun_out/sw.json "$W warps=$warps R=$R maxw=$mw" || tail -3 gpurun_out/sw.err
done; done; done; done
