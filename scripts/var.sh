#!/bin/bash
# runs the GPU parity tests + a short bench for each kernel variant (experiments only)
for v in realsensetracker_b200/_lib/variants/*.so; do
  echo "== variant: $v"
  RST_ALIGN_LIB=$v python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -1
  RST_ALIGN_LIB=$v python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'l0_us',round(d['roofline']['avg_launch_us'],1),'frac',round(d['roofline']['frac'],3),'survey',round(d['roofline']['survey_equiv']['frac'],3), d['roofline']['stage_share_of_step'])"
done
