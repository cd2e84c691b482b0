#!/usr/bin/env python
"""CPU-side models of k_lift_planes (csrc/lift_planes.cu) on the bench scene -- no GPU, no oracle: the projection is
restated in numpy (fp32; a simulation input, not a parity path).

1. Shared-memory wavefronts of the plane gathers for alternative lane -> voxel mappings and plane row pitches:
   one LDS of a warp costs max over the 32 banks of the number of DISTINCT words addressed in that bank; invalid
   voxel-views read the zero word behind the plane.  Question: is there a mapping / pitch with fewer bank conflicts?
2. Work per compute warp with the kernel's oct pairing (k-th most with k-th least expensive oct), cost model
   245 cycles per (warp, view) that sees any quad + 130 per active quad (from the per-warp clock trace).
   Question: how far is the busiest warp above the mean, i.e. what does oct-granular ownership cost?

  python tools/sim_lift_smem.py            (about a minute)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import lifting  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_scene  # noqa: E402

H, W = 59, 80
GRID = (40, 40, 16)


def project(seed, nv=50):
    cfg = SceneConfig(channels=4, n_views=nv, n_voxels=GRID)
    sc = make_scene(cfg, seed=seed, with_images=False)
    proj = lifting.compute_projection(sc.img_meta, 4).numpy().astype(np.float32)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).numpy().reshape(3, -1)
    hom = np.concatenate([pts, np.ones((1, pts.shape[1]), np.float32)], 0)
    q = np.einsum('vrk,kn->vrn', proj, hom).astype(np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        x = np.rint(q[:, 0] / q[:, 2])
        y = np.rint(q[:, 1] / q[:, 2])
    valid = (x >= 0) & (y >= 0) & (x < W) & (y < H) & (q[:, 2] > 0)
    x = np.where(valid, x, 0).astype(np.int64).reshape(nv, *GRID)
    y = np.where(valid, y, 0).astype(np.int64).reshape(nv, *GRID)
    return x, y, valid.reshape(nv, *GRID)


def mapping_block(bx, by, ov=8):
    """the kernel's compact oct: lane = (lx, ly) in a bx x by block of columns, ov consecutive z per lane."""
    ix, iy, iz, qs, q = [], [], [], [], 0
    lane = np.arange(32)
    for tx in range(GRID[0] // bx):
        for ty in range(GRID[1] // by):
            for o in range(GRID[2] // ov):
                for k in range(ov):
                    ix.append(tx * bx + lane // by); iy.append(ty * by + lane % by); iz.append(np.full(32, o * ov + k))
                    qs.append(q + k // 4)
                q += ov // 4
    return np.array(ix), np.array(iy), np.array(iz), np.array(qs)


def mapping_zcols(nz, bx, by, run):
    """lanes = nz z-groups x (bx x by) columns; every lane owns `run` consecutive z."""
    ix, iy, iz, qs, q = [], [], [], [], 0
    lane = np.arange(32)
    zt = nz * run
    for tx in range(GRID[0] // bx):
        for ty in range(GRID[1] // by):
            for o in range(GRID[2] // zt):
                for k in range(run):
                    lz, lc = lane % nz, lane // nz
                    ix.append(tx * bx + lc // by); iy.append(ty * by + lc % by); iz.append(o * zt + lz * run + k)
                    qs.append(q + k // 4)
                q += max(run // 4, 1)
    return np.array(ix), np.array(iy), np.array(iz), np.array(qs)


def wavefronts(x, y, valid, mapping, pitch):
    ix, iy, iz, quad = mapping
    tot_w = tot_i = lanes = 0
    for v in range(x.shape[0]):
        ok = valid[v][ix, iy, iz]
        addr = np.where(ok, y[v][ix, iy, iz] * pitch + x[v][ix, iy, iz], H * pitch)
        qact = np.zeros(quad.max() + 1, bool)
        np.logical_or.at(qact, quad, ok.any(1))
        act = qact[quad]                                    # the kernel skips quads no lane of which is valid
        a = addr[act]
        key = np.sort((a % 32) * 100000 + a, axis=1)
        first = np.ones_like(key, bool)
        first[:, 1:] = key[:, 1:] != key[:, :-1]
        bank = key // 100000
        w = np.zeros(len(a), int)
        for b in range(32):
            w = np.maximum(w, ((bank == b) & first).sum(1))
        tot_w += w.sum(); tot_i += len(a); lanes += ok[act].sum()
    return tot_w, tot_i, lanes


def warp_balance(valid, n_parts=2, warps=25, view_w=3):
    nv = valid.shape[0]
    v = valid.reshape(nv, 10, 4, 5, 8, 2, 2, 4)
    qm = v.any(axis=(2, 4, 7)).reshape(nv, 100, 2)          # [view, oct, quad]
    cost = qm.sum(axis=(0, 2)) + view_w * qm.any(axis=2).sum(axis=0)
    order = sorted(range(100), key=lambda q: (-cost[q], q))
    out = []
    for part in range(n_parts):
        work = []
        for w in range(warps):
            pair = w * n_parts + part
            qa, qb = order[pair], order[99 - pair]
            nq = qm[:, qa].sum(1) + qm[:, qb].sum(1)
            work.append(np.where(nq > 0, 245 + 130 * nq, 0).sum())
        out.append((np.mean(work), np.max(work)))
    hottest = max(np.where(qm[:, q].sum(1) > 0, 245 + 130 * qm[:, q].sum(1), 0).sum() for q in range(100))
    return out, hottest


def quad_balance(valid, n_warps=50, per_warp=4):
    """What quad-granular ownership would give: every warp owns 4 quads (128 voxels each) picked greedily -- the most
    expensive quad first, to the warp whose total grows least (a warp pays the 245 cycles of a view once, whichever of
    its quads see it)."""
    nv = valid.shape[0]
    qm = valid.reshape(nv, 10, 4, 5, 8, 2, 2, 4).any(axis=(2, 4, 7)).reshape(nv, 200)     # [view, quad]
    order = np.argsort(-qm.sum(0), kind='stable')
    seen = np.zeros((n_warps, nv), bool)
    nq = np.zeros((n_warps, nv), int)
    cnt = np.zeros(n_warps, int)
    for q in order:
        best, best_t = None, None
        for w in range(n_warps):
            if cnt[w] >= per_warp:
                continue
            t = (245 * (seen[w] | qm[:, q]) + 130 * (nq[w] + qm[:, q])).sum()
            if best is None or t < best_t:
                best, best_t = w, t
        seen[best] |= qm[:, q]; nq[best] += qm[:, q]; cnt[best] += 1
    tot = (245 * seen + 130 * nq).sum(1)
    return tot.mean(), tot.max()


def main():
    x, y, valid = project(0)
    print(f'valid voxel-views: {valid.mean():.3f}')
    maps = [('kernel: 4x8 columns x 8 z', mapping_block(4, 8)), ('8x4 columns x 8 z', mapping_block(8, 4)),
            ('2x16 columns x 8 z', mapping_block(2, 16)), ('2 z-groups x 4x4 columns, run 8', mapping_zcols(2, 4, 4, 8)),
            ('4 z-groups x 2x4 columns, run 4', mapping_zcols(4, 2, 4, 4)), ('16 z-groups x 1x2 columns, run 1', mapping_zcols(16, 1, 2, 1))]
    for name, m in maps:
        for pitch in (80, 84, 88, 81):
            w, i, ln = wavefronts(x, y, valid, m, pitch)
            print(f'{name:34s} row pitch {pitch}: {i / 1e3:6.1f}k LDS per channel, {w / 1e3:6.1f}k wavefronts '
                  f'({w / i:.2f} per LDS, {ln / i:.1f} valid lanes per LDS)', flush=True)
    for seed in (0, 1, 1000):
        _, _, valid = project(seed)
        parts, hottest = warp_balance(valid)
        for p, (mean, mx) in enumerate(parts):
            print(f'seed {seed} part {p}: work per warp and unit: mean {mean:.0f} cycles, busiest {mx:.0f} ({mx / mean:.2f}x); '
                  f'hottest single oct {hottest:.0f}')
        qmean, qmax = quad_balance(valid)
        print(f'seed {seed}: quad-granular greedy ownership (4 quads per warp): mean {qmean:.0f}, busiest {qmax:.0f} ({qmax / qmean:.2f}x)')


if __name__ == '__main__':
    main()
