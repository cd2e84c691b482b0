#!/bin/bash
# A/B timing of the fused lift under environment knobs: tools/lift_ab.sh "ND_LIFT_PREFETCH=0" "ND_LIFT_PREFETCH=3 ND_LIFT_DEBUG=17" ...
for cfg in "$@"; do
  r=$(env $cfg python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), round(d['roofline']['frac'],3))")
  echo "$cfg -> us_per_step, frac = $r"
done
