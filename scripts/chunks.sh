#!/bin/bash
# blocking-call / e2e bench over --chunk sizes (in-call H2D/compute pipelining experiment)
for c in 0 33 44 65; do
  python bench.py --steps 20 --warmup 3 --no-cpu --chunk $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $c value',round(d['value']),'e2e',round(d['e2e']['value']),'blocking',round(d['blocking_call']['value']),'blocking_ms',round(d['blocking_call']['ms_per_step'],3))"
done
