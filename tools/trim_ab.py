#!/usr/bin/env python
"""A/B of lift-kernel variants on one GPU, in ONE process (same buffers, alternating): 148 CTAs (ND_LIFT_NO_TRIM=1) against the
grid trimmed to the rounds the units need (128 CTAs at the bench shape), other SM limits (ND_LIFT_SMS), and the experimental
entry-driven consumer loop with prefetched offset rows (ND_LIFT_PREFETCH=1; parity is checked first)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from nerfdet_b200 import lifting  # noqa: E402


def main():
    dev = torch.device('cuda', 0)
    proj, pts = bench.build_scene(1000, bench.NV_PER_GPU)
    proj, pts = proj.to(dev), pts.to(dev)
    sets = [bench.host_features(2000 + i, bench.NV_PER_GPU).to(dev) for i in range(bench.N_INPUT_SETS)]
    h, w = bench.FEAT_HW

    def run(steps=200):
        for i in range(10):
            lifting.lift_mean_var(sets[i % 3][:, :, :h, :w], pts, proj)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            lifting.lift_mean_var(sets[i % 3][:, :, :h, :w], pts, proj)
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) / steps * 1e3, 1)

    modes = {'148 CTAs': {'ND_LIFT_NO_TRIM': '1'}, '128 CTAs (trimmed)': {}, '136 CTAs': {'ND_LIFT_NO_TRIM': '1', 'ND_LIFT_SMS': '136'},
             'prefetching consumers (experimental)': {'ND_LIFT_PREFETCH': '1'}}
    knobs = ('ND_LIFT_NO_TRIM', 'ND_LIFT_SMS', 'ND_LIFT_PREFETCH')
    # parity of the experimental consumer loop against the default one before it is timed
    ref = lifting.lift_mean_var(sets[0][:, :, :h, :w], pts, proj)
    os.environ['ND_LIFT_PREFETCH'] = '1'
    exp = lifting.lift_mean_var(sets[0][:, :, :h, :w], pts, proj)
    os.environ.pop('ND_LIFT_PREFETCH')
    torch.cuda.synchronize()
    print(f'prefetching consumers vs default: counts equal {bool(torch.equal(ref[2], exp[2]))}, mean max abs diff '
          f'{float((ref[0] - exp[0]).abs().max()):.3e}, cov max abs diff {float((ref[1] - exp[1]).abs().max()):.3e} '
          f'(same view order, so 0 is expected)', flush=True)
    res = {k: [] for k in modes}
    for rep in range(5):
        for name, env in modes.items():
            for k in knobs:
                os.environ.pop(k, None)
            os.environ.update(env)
            res[name].append(run())
    for k in knobs:
        os.environ.pop(k, None)
    for name, v in res.items():
        print(f'{name:22s} us/step over 5 alternating repeats: {v}  median {sorted(v)[2]}', flush=True)


if __name__ == '__main__':
    main()
