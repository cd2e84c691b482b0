#!/bin/bash
# Sweep the plane-resident lift's tuning knobs on the GPU box: prints us per step per configuration.
# usage: tools/sweep_lift.sh "S W DEBUG" ...
for cfg in "$@"; do
  set -- $cfg
  r=$(ND_LIFT_STAGES=$1 ND_LIFT_WARPS=$2 ND_LIFT_DEBUG=${3:-0} python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1))")
  echo "stages=$1 warps=$2 debug=${3:-0} us_per_step=$r"
done
