#!/usr/bin/env python
"""Multi-GPU check of the view-sharded lift, run under torchrun with one rank per GPU:
every rank lifts its slice of the views, the accumulators are summed with one NCCL all-reduce, and the finalised
mean / all-view variance / count are compared with the single-GPU fused lift over ALL views run on the same device.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfdet_b200 import distributed as nd_dist  # noqa: E402
from nerfdet_b200 import lifting  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_scene  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    nv = 100                                              # BASELINE configs[3]: ~100 views at test time
    cfg = SceneConfig(n_views=nv, n_voxels=(40, 40, 16), voxel_size=(.16, .16, .2), channels=256)
    sc = make_scene(cfg, seed=77, with_images=False, with_features=True)      # same seed -> same scene on every rank
    proj = lifting.compute_projection(sc.img_meta, 4).to(dev)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin']).to(dev)
    feats = sc.features.to(dev)[:, :, :59, :80]
    b, e = nd_dist.view_shard(nv, rank, world)
    mean, cov, cnt = nd_dist.lift_mean_var_view_sharded(feats[b:e], pts, proj[b:e], n_views_total=nv)
    m1, c1, n1 = lifting.lift_mean_var(feats, pts, proj)
    torch.cuda.synchronize()
    ok_cnt = torch.equal(cnt, n1)

    def err(a, r):
        a, r = a.double(), r.double()
        tol = 1e-4 * r.abs() + 1e-5 * r.abs().max()
        return float((a - r).abs().max()), int(((a - r).abs() > tol).sum())
    e_mean, bad_mean = err(mean, m1)
    e_cov, bad_cov = err(cov, c1)
    res = torch.tensor([int(ok_cnt), bad_mean, bad_cov], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(f'view-sharded lift on {world} GPUs vs single-GPU lift over {nv} views: counts equal on {int(res[0])}/{world} ranks, '
              f'mean max abs err {e_mean:.3e} ({int(res[1])} outside 1e-4), cov max abs err {e_cov:.3e} ({int(res[2])} outside 1e-4)',
              flush=True)
    assert int(res[0]) == world and int(res[1]) == 0 and int(res[2]) == 0

    # ---- the same exchange over NVLink peer memory (csrc/peer.cu): CUDA IPC segments, epoch flags, no NCCL on the data path ----
    peer = nd_dist.PeerLift(256, pts[0].numel(), dev)
    for _ in range(3):
        pm, pc, pn = peer(feats[b:e], pts, proj[b:e], nv)
    peer.check()
    e_mean, bad_mean = err(pm, m1)
    e_cov, bad_cov = err(pc, c1)
    same = torch.equal(pm, mean) and torch.equal(pc, cov)          # identical to the all-reduce path? (not required)
    res = torch.tensor([int(torch.equal(pn, n1)), bad_mean, bad_cov, int(same)], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.SUM)
    # every rank must hold the same bits: checksum of the outputs
    chk = torch.stack([pm.double().sum(), pc.double().sum()])
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)

    def timed(fn, steps=50):
        for _ in range(5):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    us_peer = timed(lambda: peer(feats[b:e], pts, proj[b:e], nv))
    us_xchg = timed(lambda: peer.exchange(nv))
    us_nccl = timed(lambda: nd_dist.lift_mean_var_view_sharded(feats[b:e], pts, proj[b:e], n_views_total=nv))
    us_local = timed(lambda: lifting.lift_mean_var(feats[b:e], pts, proj[b:e]))
    sweep = {}
    fast = os.environ.get('DIST_CHECK_FAST') == '1'            # skip the grid sweeps (8-GPU runs are charged 8x)
    for per_sm in (() if fast else (2, 4, 6, 8, 12, 16)):      # grid of the exchange kernel (loads in flight over the links)
        os.environ['ND_PEER_CTAS_PER_SM'] = str(per_sm)
        sweep[per_sm] = round(timed(lambda: peer.exchange(nv)), 1)
    os.environ.pop('ND_PEER_CTAS_PER_SM', None)
    narrow = {}
    for ctas in (8, 12, 16, 20, 32, 48):                       # narrow grid: that many one-per-SM CTAs (PeerLift(overlap_sms=...))
        peer.overlap_sms = ctas
        narrow[ctas] = round(timed(lambda: peer.exchange(nv)), 1)
    peer.overlap_sms = 0
    peer.check()
    if rank == 0:
        print(f'exchange alone vs CTAs per SM: {sweep}', flush=True)
        print(f'exchange alone, narrow grid, vs number of fat CTAs: {narrow}', flush=True)
        print(f'peer-memory exchange on {world} GPUs: counts equal on {int(res[0])}/{world} ranks, mean max abs err {e_mean:.3e} '
              f'({int(res[1])} outside 1e-4), cov max abs err {e_cov:.3e} ({int(res[2])} outside 1e-4), bit-identical to the '
              f'all-reduce path on {int(res[3])}/{world} ranks, checksums equal across ranks: {bool(torch.equal(lo_, hi_))}', flush=True)
        print(f'per step ({nv // world} views per GPU, max over ranks): lift + peer exchange {us_peer:.1f} us '
              f'(exchange alone {us_xchg:.1f} us), lift + NCCL all-reduce + finalise {us_nccl:.1f} us, '
              f'local fused lift alone {us_local:.1f} us', flush=True)
    assert int(res[0]) == world and int(res[1]) == 0 and int(res[2]) == 0 and torch.equal(lo_, hi_)
    peer.close()

    # ---- the same kernel over NVLS multicast: sums taken in the switch (multimem.ld_reduce), rows broadcast (multimem.st) ----
    mcp, why = None, 'skipped (DIST_CHECK_NO_MC=1)'
    try:
        if os.environ.get('DIST_CHECK_NO_MC') != '1':
            mcp = nd_dist.PeerLift(256, pts[0].numel(), dev, transport='multicast')
    except Exception as ex:
        why = f'{type(ex).__name__}: {ex}'
    ok = torch.tensor([1 if mcp is not None else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok) == 0:
        if rank == 0:
            print(f'multicast transport unavailable on this box: {why}', flush=True)
    else:
        for _ in range(3):
            qm, qc, qn = mcp(feats[b:e], pts, proj[b:e], nv)
        mcp.check()
        e_mean, bad_mean = err(qm, m1)
        e_cov, bad_cov = err(qc, c1)
        res = torch.tensor([int(torch.equal(qn, n1)), bad_mean, bad_cov], device=dev)
        dist.all_reduce(res, op=dist.ReduceOp.SUM)
        chk = torch.stack([qm.double().sum(), qc.double().sum()])
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        us_mc = timed(lambda: mcp(feats[b:e], pts, proj[b:e], nv))
        sweep = {}
        for per_sm in ((0, 2, 4) if fast else (0, 2, 4, 6, 8, 12)):
            if per_sm:
                os.environ['ND_PEER_CTAS_PER_SM'] = str(per_sm)
            sweep[per_sm] = round(timed(lambda: mcp.exchange(nv)), 1)
        os.environ.pop('ND_PEER_CTAS_PER_SM', None)
        mcp.check()
        if rank == 0:
            print(f'multicast exchange on {world} GPUs: counts equal on {int(res[0])}/{world} ranks, mean max abs err {e_mean:.3e} '
                  f'({int(res[1])} outside 1e-4), cov max abs err {e_cov:.3e} ({int(res[2])} outside 1e-4), checksums equal '
                  f'across ranks: {bool(torch.equal(lo_, hi_))}', flush=True)
            print(f'per step: lift + multicast exchange {us_mc:.1f} us; exchange alone vs CTAs per SM (0 = default): {sweep}', flush=True)
        assert int(res[0]) == world and int(res[1]) == 0 and int(res[2]) == 0 and torch.equal(lo_, hi_)
        mcp.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
