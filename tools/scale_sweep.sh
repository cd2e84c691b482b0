#!/bin/bash
# N-GPU tuning sweep of the pipelined exchange: tools/scale_sweep.sh N "36 44" "2 3"
N=$1; SMS=${2:-"28 36 44"}; LANES=${3:-"2"}
for s in $SMS; do for l in $LANES; do
  out=gpurun_out/sweep_n${N}_sms${s}_lanes${l}
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
    bench.py --gpus $N --steps 200 --warmup 5 --no-e2e --overlap-sms $s --lanes $l > $out.json 2> $out.err
  python -c "import json,sys; d=json.loads(open('$out.json').read().strip().splitlines()[-1]); print('N=$N sms=$s lanes=$l', round(d['ms_per_step']*1e3,1), 'us/step', d.get('parity_ok'))" 2>/dev/null || tail -2 $out.err
done; done
