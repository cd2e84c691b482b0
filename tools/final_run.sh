#!/bin/bash
# everything the round's numbers come from, on one GPU: tests, smoke, the three bench workloads, the reference arm
set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r2_final_lift.json 2> gpurun_out/r2_final_lift.err; tail -c 600 gpurun_out/r2_final_lift.err
python bench.py --workload render --steps 300 > gpurun_out/r2_final_render.json 2> gpurun_out/r2_final_render.err
python bench.py --workload sweep --steps 20 > gpurun_out/r2_final_sweep.json 2> gpurun_out/r2_final_sweep.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err
python bench.py --scaling strong --views 100 --steps 300 --no-cpu-baseline > gpurun_out/r2_final_strong1.json 2> gpurun_out/r2_final_strong1.err
python tools/lift_dev.py --quick 2>&1 | tail -8 > gpurun_out/r2_final_liftdev.log
python tools/mlp_bench.py 2>&1 | tail -8 > gpurun_out/r2_final_mlp.log
ls -la gpurun_out/r2_final_*
