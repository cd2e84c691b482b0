#!/bin/bash
# bench.py at N GPUs in its lift variants (weak / strong scaling, with and without the variance volume):
#   tools/scale_run.sh N [steps] [extra bench args]        -> gpurun_out/scale_<variant>_nN<tag>.json
N=$1; K=${2:-100}; shift; shift
TAG=$(echo "$*" | tr -d ' -')
run() {  # name, extra args
  name=$1; shift
  out=gpurun_out/scale_${name}_n$N$TAG
  if [ "$N" = 1 ]; then
    python bench.py --gpus 1 --steps $K --warmup 5 --no-cpu-baseline "$@" > $out.json 2> $out.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps $K --warmup 5 "$@" > $out.json 2> $out.err
  fi
  echo "$name N=$N $TAG rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" $out.err | tail -3
}
run weak "$@"
run weak_nocov --no-cov "$@"
run strong --scaling strong --views 100 "$@"
run strong_nocov --scaling strong --views 100 --no-cov "$@"
