#!/bin/bash
# same-box A/B of two builds of libskb.so
for v in new old new old; do
  cp tools/ab/libskb_$v.so scikit-optimize_b200/csrc/libskb.so
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-ask | tail -1 > /tmp/b.json
  python -c "
import json; d=json.load(open('/tmp/b.json')); r=d['roofline']; print('$v', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],2), 'ms k2o', round(r['avg_launch_ms'],3), 'ms', d['clocks']['sm_mhz'], 'MHz', d['clocks']['reasons'], 'fp64', round(d['fp64_dmma_path']['value']/1e6,2))"
done
