#!/bin/bash
# runs a short bench for each kernel variant in _lib/variants (experiments only); prints value, level-0 launch, pre-processing ms
for v in default $(ls realsensetracker_b200/_lib/variants/*.so 2>/dev/null); do
  if [ "$v" = default ]; then unset RST_ALIGN_LIB; else export RST_ALIGN_LIB=$v; fi
  python bench.py --steps 30 --warmup 5 --no-cpu --quick 2>/dev/null | python -c "
import json,sys,re; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
ms=float(re.search(r'ms_per_step ([0-9.]+)', r['measured_in']).group(1))
print('$v','value',round(d['value']),'e2e',round(d['e2e']['value']),'l0_us',round(r['avg_launch_us'],1),'frac',round(r['frac'],3),'pre_ms',round(ms*r['stage_share_of_step']['preprocess'],4), 'step1s_ms', ms)"
done
