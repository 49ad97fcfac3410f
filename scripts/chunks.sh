#!/bin/bash
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
for c in 0 8 16 32 64; do
  python bench.py --steps 20 --warmup 3 --no-cpu --chunk $c 2>&1 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c value',round(d['value']),'e2e',round(d['e2e']['value']),'e2e_ms',round(d['e2e']['ms_per_step'],3),'l0_us',round(d['roofline']['avg_launch_us'],1))"
done
