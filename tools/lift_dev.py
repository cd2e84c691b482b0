#!/usr/bin/env python
"""Development harness of the fused lift (csrc/lift_quads.cu): parity against the C oracle, then timings of the
plan-based kernel under its tuning knobs (views per stage, stages) and of the geometry pass.

  python tools/lift_dev.py [--quick]      (GPU box; prints one line per measurement)
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import _lib, lifting, ops  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_features, make_scene  # noqa: E402

DEV = 'cuda'


def scene(nv, grid, vsize, seed):
    cfg = SceneConfig(n_views=nv, n_voxels=grid, voxel_size=vsize, channels=4)
    sc = make_scene(cfg, seed=seed, with_images=False, with_features=False)
    proj = lifting.compute_projection(sc.img_meta, 4)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin'])
    return proj, pts


def close(a, b, rtol=1e-4):
    a = a.detach().cpu().double().numpy()
    b = np.asarray(b, dtype=np.float64)
    atol = 1e-5 * max(float(np.abs(b).max()), 1e-30)
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    return int(bad.sum()), float(np.abs(a - b).max())


def parity(nv, c, grid, vsize, seed, **kw):
    from oracle import c_oracle
    proj, pts = scene(nv, grid, vsize, seed)
    rs = np.random.RandomState(seed)
    feats = torch.from_numpy(make_features(rs, (nv, c, 60, 80)))
    m_ref, c_ref, n_ref = c_oracle.lift(feats[:, :, :59, :80].numpy(), pts.numpy(), proj.numpy())
    fd, pd, qd = feats.to(DEV), pts.to(DEV), proj.to(DEV)
    plan = ops.LiftPlan(fd[:, :, :59, :80], pd, qd, **kw)
    assert plan.eligible
    out = []
    for rep in range(3):                                    # repeated launches on one plan (ticket counter)
        mean, cov, cnt = plan.mean_var(fd[:, :, :59, :80])
        torch.cuda.synchronize()
        cnt_ok = bool(np.array_equal(cnt.cpu().numpy(), n_ref))
        bm, em = close(mean, m_ref)
        bc, ec = close(cov, c_ref)
        mean_biteq = bool(np.array_equal(mean.cpu().numpy(), m_ref))
        out.append((cnt_ok, bm, bc, em, ec, mean_biteq))
    print(f'parity nv={nv} C={c} grid={grid} {kw}: ' + ' | '.join(
        f'count_ok={o[0]} mean_bad={o[1]} cov_bad={o[2]} max_err=({o[3]:.2e},{o[4]:.2e}) mean_bit_equal={o[5]}' for o in out),
        flush=True)
    return all(o[0] and o[1] == 0 and o[2] == 0 for o in out)


def timed(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3


def per_step(fn, steps, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    t = np.array([evs[i].elapsed_time(evs[i + 1]) * 1e3 for i in range(steps)])
    return float(t.min()), float(np.median(t)), float(np.percentile(t, 90))


def grids():
    nv, c = 50, 256
    sets = [torch.from_numpy(make_features(np.random.RandomState(2000 + i), (nv, c, 60, 80))).to(DEV) for i in range(3)]
    views = [s[:, :, :59, :80] for s in sets]
    for grid, vs in (((40, 40, 16), (.16, .16, .2)), ((56, 56, 16), (.16, .16, .2)), ((64, 64, 24), (.1, .1, .13)),
                     ((80, 80, 32), (.08, .08, .08))):
        n_vox = int(np.prod(grid))
        bytes_step = nv * c * 59 * 80 * 4 + 2 * c * n_vox * 4 + n_vox * 8 + nv * 48
        plans = []
        for seed in range(8):
            proj, pts = scene(nv, grid, vs, 1000 + seed)
            plans.append(ops.LiftPlan(views[0], pts.to(DEV), proj.to(DEV)))
        it = [0]

        def one():
            it[0] += 1
            return plans[0].mean_var(views[it[0] % 3])

        def rot():
            it[0] += 1
            return plans[it[0] % 8].mean_var(views[it[0] % 3])
        us1 = timed(one, 30)
        mn, med, p90 = per_step(one, 20)
        us3 = timed(rot, 32)
        each = []
        for k in range(8):
            each.append(timed(lambda: plans[k].mean_var(views[k % 3]), 6, warmup=2))
        print(f'grid {grid}: one plan back-to-back {us1:.1f} us ({bytes_step / us1 / 1e3 / 6551:.3f} of peak), event-separated '
              f'{mn:.1f}/{med:.1f}/{p90:.1f}; eight plans rotating {us3:.1f} us; each plan: {[round(e) for e in each]}; plan bytes {plans[0].bytes / 1e6:.1f} MB', flush=True)


def backward():
    nv, c, grid = 50, 256, (40, 40, 16)
    proj, pts = scene(nv, grid, (0.16, 0.16, 0.2), 1000)
    pd, qd = pts.to(DEV), proj.to(DEV)
    for dtype in (torch.float32, torch.bfloat16):
        f = torch.from_numpy(make_features(np.random.RandomState(2000), (nv, c, 60, 80))).to(DEV).to(dtype)[:, :, :59, :80]
        mean, cov, cnt = ops.lift_mean_var_planned(f, pd, qd, None, True)
        gm, gc_ = torch.randn_like(mean), torch.randn_like(cov)
        us = timed(lambda: ops.direct.lift_backward(f, pd, qd, mean, cov, cnt, gm, gc_, None, 0.0, 0), 50)
        us_m = timed(lambda: ops.direct.lift_backward(f, pd, qd, mean, None, cnt, gm, None, None, 0.0, 0), 50)
        elt = f.element_size()
        moved = 2 * nv * c * 59 * 80 * elt + 4 * c * 25600 * 4          # features in + gradient out, mean/cov/g_mean/g_cov in
        print(f'lift backward {dtype}: {us:.1f} us ({moved / us / 1e3:.0f} GB/s of features + gradient + volumes), '
              f'mean-only {us_m:.1f} us; forward for comparison '
              f'{timed(lambda: ops.lift_mean_var_planned(f, pd, qd, None, True), 100):.1f} us', flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--grids', action='store_true', help='time the other voxel grids of the sweep instead')
    ap.add_argument('--backward', action='store_true', help='time the backward of the lift at the bench shape instead')
    args = ap.parse_args()
    torch.cuda.set_device(0)
    if args.grids:
        return grids()
    if args.backward:
        return backward()
    ok = True
    ok &= parity(6, 64, (20, 20, 8), (0.32, 0.32, 0.4), 7)
    ok &= parity(3, 40, (40, 40, 16), (0.16, 0.16, 0.2), 13)
    ok &= parity(50, 32, (40, 40, 16), (0.16, 0.16, 0.2), 12)
    ok &= parity(100, 16, (40, 40, 16), (0.16, 0.16, 0.2), 14)
    ok &= parity(50, 16, (40, 40, 16), (0.16, 0.16, 0.2), 12, views_per_stage=1)
    ok &= parity(50, 16, (40, 40, 16), (0.16, 0.16, 0.2), 12, stages=2)
    ok &= parity(20, 16, (56, 56, 16), (0.16, 0.16, 0.2), 15)
    ok &= parity(7, 8, (13, 11, 7), (0.4, 0.4, 0.4), 16)            # linear tiling, ragged quads
    print('PARITY', 'OK' if ok else 'FAILED', flush=True)

    # ---- timings at the bench shape ----
    nv, c, grid = 50, 256, (40, 40, 16)
    proj, pts = scene(nv, grid, (0.16, 0.16, 0.2), 1000)
    pd, qd = pts.to(DEV), proj.to(DEV)
    sets = [torch.from_numpy(make_features(np.random.RandomState(2000 + i), (nv, c, 60, 80))).to(DEV) for i in range(3)]
    views = [s[:, :, :59, :80] for s in sets]
    n_vox = int(np.prod(grid))
    bytes_step = nv * c * 59 * 80 * 4 + 2 * c * n_vox * 4 + n_vox * 8 + nv * 48
    variants = [dict(), dict(stages=2), dict(stages=3), dict(stages=4), dict(stages=5), dict(stages=6), dict(views_per_stage=1),
                dict(views_per_stage=1, stages=6), dict(prefetch_stages=6)]
    if args.quick:
        variants = variants[:1]
    for kw in variants:
        plan = ops.LiftPlan(views[0], pd, qd, **kw)
        it = [0]

        def step():
            it[0] += 1
            return plan.mean_var(views[it[0] % 3])
        us = timed(step, args.steps)
        mn, med, p90 = per_step(step, 60)
        print(f'quads {kw}: back-to-back {us:.1f} us/step ({bytes_step / us / 1e3:.0f} GB/s, frac {bytes_step / us / 1e3 / 6551:.3f}); '
              f'event-separated min/median/p90 {mn:.1f}/{med:.1f}/{p90:.1f} us', flush=True)
    # geometry pass alone and the one-shot op (plan + lift per call), and the cached Python API
    it = [0]

    def build_only():
        ops.LiftPlan(views[0], pd, qd)
    print(f'geometry plan build: {timed(build_only, 100):.1f} us', flush=True)

    def oneshot():
        it[0] += 1
        return ops.lift_mean_var(views[it[0] % 3], pd, qd, None, True, 0)
    print(f'one-shot op (plan + lift, torch.library dispatch): {timed(oneshot, args.steps):.1f} us/step', flush=True)

    def api():
        it[0] += 1
        return lifting.lift_mean_var(views[it[0] % 3], pd, qd)
    print(f'lifting.lift_mean_var (cached plan): {timed(api, args.steps):.1f} us/step', flush=True)
    t0 = time.perf_counter()
    for _ in range(2000):
        api()
    host_us = (time.perf_counter() - t0) / 2000 * 1e6
    torch.cuda.synchronize()
    print(f'host time per lifting.lift_mean_var call (enqueue only): {host_us:.1f} us', flush=True)


if __name__ == '__main__':
    main()
