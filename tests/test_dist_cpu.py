"""CPU tier for the multi-GPU path: the view-sharded lift's host logic -- view partition,
one all-reduce of the flat [S1 | S2 | count] accumulators, finalise with the GLOBAL view count -- on two ``gloo``
ranks.  The CUDA ops are replaced by oracle-backed stand-ins with the same buffer contract (this file is test
infrastructure: the product path has no CPU fallback), and the result is compared with the oracle's lift over ALL
views on one process."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from nerfdet_b200 import distributed as nd_dist  # noqa: E402
from nerfdet_b200 import lifting  # noqa: E402
from nerfdet_b200.synthetic import SceneConfig, make_scene  # noqa: E402
from oracle import lift_oracle as lo  # noqa: E402


def oracle_accumulate(features, points, projection):
    """Stand-in for ops.lift_accumulate: [S1 (C*N) | S2 (C*N) | count (N)] float32 over this rank's views."""
    volume, valid = lo.backproject(features, points, projection, None, None)       # [nv, C, X, Y, Z], [nv, 1, X, Y, Z]
    nv, c = volume.shape[:2]
    v = volume.reshape(nv, c, -1)
    s1 = v.sum(dim=0)
    s2 = (v * v).sum(dim=0)
    cnt = valid.reshape(nv, -1).sum(dim=0).to(torch.float32)
    return torch.cat([s1.reshape(-1), s2.reshape(-1), cnt])


def oracle_finalize(acc, n_views_total, channels, n_voxels, alpha, want_cov):
    """Stand-in for ops.lift_finalize (nerfdet.py:171-181 restated from accumulators)."""
    c, n = channels, n_voxels
    s1, s2, cnt = acc[:c * n].view(c, n), acc[c * n:2 * c * n].view(c, n), acc[2 * c * n:]
    seen = cnt > 0
    denom = torch.where(seen, cnt, torch.ones_like(cnt))
    mean = torch.where(seen, s1 / denom, torch.zeros_like(s1))
    ssd = s2 - 2 * mean * s1 + float(n_views_total) * mean * mean                   # all views, invalid ones included
    cov = torch.where(seen, torch.exp(-(ssd.clamp_min(0) / denom)), torch.zeros_like(s1))
    if alpha is not None:
        mean = mean * alpha.view(1, n)
    return mean, cov if want_cov else mean.new_empty(0), cnt.to(torch.int64)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nv, pass_total, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = SceneConfig(n_views=nv, n_voxels=(8, 8, 4), voxel_size=(0.8, 0.8, 0.8), channels=6)
        sc = make_scene(cfg, seed=21, with_images=False)                           # same seed: same scene on both ranks
        proj = lifting.compute_projection(sc.img_meta, 4)
        pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin'])
        feats = sc.features[:, :, :59, :80]
        b, e = nd_dist.view_shard(nv, rank, world)
        mean, cov, cnt = nd_dist.lift_mean_var_view_sharded(
            feats[b:e], pts, proj[b:e], n_views_total=nv if pass_total else None,
            accumulate_fn=oracle_accumulate, finalize_fn=oracle_finalize)
        m_ref, c_ref, n_ref = lo.lift_mean_var(feats, pts, proj)
        ok_cnt = bool(torch.equal(cnt.view(-1), n_ref.view(-1)))
        e_mean = float((mean - m_ref.view_as(mean)).abs().max())
        e_cov = float((cov - c_ref.view_as(cov)).abs().max())
        queue.put((rank, e - b, ok_cnt, e_mean, e_cov, float(m_ref.abs().max())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('nv,pass_total', [(6, True), (5, False)])
def test_view_sharded_lift_two_gloo_ranks(nv, pass_total):
    """Even and uneven view splits; the global view count passed in or all-reduced."""
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nv, pass_total, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[1] for r in results) == sorted([nv // 2, nv - nv // 2])
    for rank, _, ok_cnt, e_mean, e_cov, scale in results:
        assert ok_cnt, f'rank {rank}: counts differ'
        assert e_mean <= 1e-5 * max(scale, 1.0) and e_cov <= 1e-5, (rank, e_mean, e_cov)


def test_view_shard_partitions():
    for nv in (1, 2, 5, 50, 100, 101):
        for world in (1, 2, 3, 4, 8):
            cuts = [nd_dist.view_shard(nv, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == nv
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        nd_dist.view_shard(4, 2, 2)


def test_channel_shard_partitions_channels():
    """The channel slices of the peer-memory exchange (csrc/peer.cu computes the same split) tile [0, C) exactly."""
    for c in (1, 3, 32, 35, 256):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = nd_dist.channel_shard(c, r, world)
                assert b == prev and e >= b and e - b in (c // world, c // world + 1)
                prev = e
            assert prev == c


# ---- the peer-memory exchange (csrc/peer.cu) as data movement: "segments every rank can read" = all_gather of the flat
# accumulators, rank r reduces + finalises ONLY its channel slice in rank order, "stores into every rank" = all_gather of the
# finished slices.  The CUDA kernel itself is tested on the GPU (tests/test_peer_gpu.py, tools/dist_check.py); this checks
# that slice ownership + per-slice finalisation reproduces the lift over all views and leaves identical bits on every rank.
def _peer_worker(rank, world, port, nv, channels, queue):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        cfg = SceneConfig(n_views=nv, n_voxels=(8, 8, 4), voxel_size=(0.8, 0.8, 0.8), channels=channels)
        sc = make_scene(cfg, seed=23, with_images=False)
        proj = lifting.compute_projection(sc.img_meta, 4)
        pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin'])
        feats = sc.features[:, :, :59, :80]
        n = 8 * 8 * 4
        b, e = nd_dist.view_shard(nv, rank, world)
        acc = oracle_accumulate(feats[b:e], pts, proj[b:e])
        segments = [torch.empty_like(acc) for _ in range(world)]
        dist.all_gather(segments, acc)                                            # peer-readable segments
        c0, c1 = nd_dist.channel_shard(channels, rank, world)
        s1 = sum(s[:channels * n].view(channels, n)[c0:c1] for s in segments)     # rank order, like the kernel
        s2 = sum(s[channels * n:2 * channels * n].view(channels, n)[c0:c1] for s in segments)
        cnt = sum(s[2 * channels * n:] for s in segments)
        m_slice, c_slice, count = oracle_finalize(torch.cat([s1.reshape(-1), s2.reshape(-1), cnt]), nv, c1 - c0, n, None, True)
        # rows stored into every rank's outputs: slices may have different sizes -> pad to the largest
        width = -(-channels // world)
        send = torch.zeros(2, width, n)
        send[0, :c1 - c0], send[1, :c1 - c0] = m_slice, c_slice
        recv = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(recv, send)
        mean = torch.cat([recv[r][0, :nd_dist.channel_shard(channels, r, world)[1] - nd_dist.channel_shard(channels, r, world)[0]]
                          for r in range(world)])
        cov = torch.cat([recv[r][1, :nd_dist.channel_shard(channels, r, world)[1] - nd_dist.channel_shard(channels, r, world)[0]]
                         for r in range(world)])
        m_ref, c_ref, n_ref = lo.lift_mean_var(feats, pts, proj)
        chk = torch.stack([mean.double().sum(), cov.double().sum()])
        lo_, hi_ = chk.clone(), chk.clone()
        dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
        queue.put((rank, bool(torch.equal(count, n_ref.view(-1))), float((mean - m_ref.view(channels, n)).abs().max()),
                   float((cov - c_ref.view(channels, n)).abs().max()), float(m_ref.abs().max()), bool(torch.equal(lo_, hi_))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('nv,channels', [(6, 6), (5, 7)])
def test_peer_exchange_data_movement_two_gloo_ranks(nv, channels):
    ctx = mp.get_context('spawn')
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, nv, channels, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_cnt, e_mean, e_cov, scale, same in results:
        assert ok_cnt and same, f'rank {rank}: counts differ or ranks hold different results'
        assert e_mean <= 1e-5 * max(scale, 1.0) and e_cov <= 1e-5, (rank, e_mean, e_cov)
