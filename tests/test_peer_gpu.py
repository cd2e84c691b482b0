"""GPU tier: the peer-memory exchange step of the view-sharded lift (csrc/peer.cu, distributed.PeerLift) on ONE GPU.

(i) world = 1: accumulate into the segment + nd_lift_finalize_peers must equal the fused lift (full benchmark shape,
    and a voxel count that is not a multiple of 4 -> scalar kernel);
(ii) the multi-rank protocol (epoch flags, channel slices, P2P loads / stores, completion counter) with 2, 3 and 8
     in-process ranks whose segments live on the same device, each rank on its own stream, over several epochs, against
     the single-shard lift over all views and against the NCCL-path finalise (lift_accumulate -> sum -> lift_finalize).
The same protocol across real GPUs (CUDA IPC segments) is exercised by tools/dist_check.py under torchrun.

Tolerances: counts bit-exact; floats |a-b| <= 1e-4*|b| + 1e-5*max|b| (sums over shards are re-associated)."""
import numpy as np
import pytest
import torch

from nerfdet_b200 import distributed as nd_dist
from nerfdet_b200 import lifting, ops
from nerfdet_b200.synthetic import SceneConfig, make_scene

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def assert_close(a, b, rtol, name=''):
    a = a.detach().cpu().double().numpy()
    b = b.detach().cpu().double().numpy()
    assert a.shape == b.shape, (name, a.shape, b.shape)
    atol = 1e-5 * max(float(np.abs(b).max()), 1e-30)
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    assert not bad.any(), f'{name}: {bad.sum()}/{bad.size} outside tolerance, max abs err {np.abs(a - b).max():.3e}'


def _scene(nv, n_voxels, voxel_size, channels, seed):
    cfg = SceneConfig(n_views=nv, n_voxels=n_voxels, voxel_size=voxel_size, channels=channels)
    sc = make_scene(cfg, seed=seed, with_images=False)
    proj = lifting.compute_projection(sc.img_meta, 4)
    pts = lifting.get_points(cfg.n_voxels, cfg.voxel_size, sc.img_meta['lidar2img']['origin'])
    return sc.features.to(DEV)[:, :, :59, :80], pts.to(DEV), proj.to(DEV)


@pytest.mark.parametrize('nv,grid,channels,overlap', [(50, (40, 40, 16), 256, 0), (50, (40, 40, 16), 256, 20),
                                                      (6, (7, 5, 3), 16, 0), (6, (7, 5, 3), 16, 4)])
def test_single_rank_equals_fused_lift(nv, grid, channels, overlap):
    """overlap > 0: accumulate with an SM limit + the narrow (few fat CTAs) exchange grid."""
    f, pts, proj = _scene(nv, grid, (0.16, 0.16, 0.2), channels, 91)
    mean, cov, cnt = lifting.lift_mean_var(f, pts, proj)
    peer = nd_dist.PeerLift(channels, int(np.prod(grid)), DEV, overlap_sms=overlap)
    try:
        for _ in range(2):                                  # two epochs through the same segment
            m2, c2, n2 = peer(f, pts, proj, nv)
        peer.check()
        assert torch.equal(n2, cnt)
        assert_close(m2, mean, 1e-5, 'mean')
        assert_close(c2, cov, 1e-4, 'cov')
    finally:
        peer.close()


@pytest.mark.parametrize('world,overlap', [(2, 0), (3, 0), (8, 0), (2, 3), (4, 2), (8, 2)])
def test_in_process_ranks_match_all_views(world, overlap):
    nv, grid, channels = 19, (16, 16, 8), 20                # uneven view split, channel slices of unequal size
    n = int(np.prod(grid))
    f, pts, proj = _scene(nv, grid, (0.4, 0.4, 0.4), channels, 92)
    alpha = torch.rand(n, device=DEV)
    mean, cov, cnt = lifting.lift_mean_var(f, pts, proj, alpha=alpha)
    # the NCCL-path arithmetic on the same shards
    acc = None
    for r in range(world):
        b, e = nd_dist.view_shard(nv, r, world)
        a = ops.lift_accumulate(f[b:e], pts, proj[b:e], 0)
        acc = a if acc is None else acc + a
    m_ref, c_ref, n_ref = ops.lift_finalize(acc, nv, channels, n, alpha, True)
    ranks = nd_dist.PeerLift.local_group(world, channels, n, DEV, overlap_sms=overlap)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        torch.cuda.synchronize()
        for epoch in range(3):
            outs = []
            for r, (peer, st) in enumerate(zip(ranks, streams)):
                b, e = nd_dist.view_shard(nv, r, world)
                with torch.cuda.stream(st):
                    outs.append(peer(f[b:e], pts, proj[b:e], nv, alpha=alpha))
            torch.cuda.synchronize()
        for peer in ranks:
            peer.check()
        for r, (m2, c2, n2) in enumerate(outs):
            assert torch.equal(n2, cnt), f'count of rank {r}'
            assert torch.equal(n2.view(-1), n_ref)
            assert_close(m2, mean, 1e-4, f'mean of rank {r}')
            assert_close(c2, cov, 1e-4, f'cov of rank {r}')
            assert_close(m2.reshape(channels, -1), m_ref, 1e-5, f'mean of rank {r} vs all-reduce path')
            assert_close(c2.reshape(channels, -1), c_ref, 1e-5, f'cov of rank {r} vs all-reduce path')
            # every rank holds the same bits: each row was computed once and stored into every segment
            assert torch.equal(m2, outs[0][0]) and torch.equal(c2, outs[0][1])
        # and the exchange against the CPU oracle lifting the whole scene (its mean times alpha, like nerfdet.py:259-261)
        from oracle import c_oracle
        m_or, c_or, n_or = c_oracle.lift(f.cpu().numpy(), pts.cpu().numpy(), proj.cpu().numpy())
        assert np.array_equal(outs[0][2].view(-1).cpu().numpy(), n_or)
        assert_close(outs[0][0].reshape(channels, -1), torch.from_numpy(m_or * alpha.cpu().numpy()[None, :]), 1e-4, 'mean vs the C oracle')
        assert_close(outs[0][1].reshape(channels, -1), torch.from_numpy(c_or), 1e-4, 'cov vs the C oracle')
    finally:
        for peer in ranks:
            peer.close()


@pytest.mark.parametrize('world,overlap', [(2, 0), (4, 2), (8, 0)])
def test_owner_rank_receives_the_scene(world, overlap):
    """owner >= 0: only that rank's outputs are written (bit-identical to the all-gather form), the owner rotates with the scene."""
    nv, grid, channels = 19, (16, 16, 8), 20
    n = int(np.prod(grid))
    f, pts, proj = _scene(nv, grid, (0.4, 0.4, 0.4), channels, 95)
    ranks = nd_dist.PeerLift.local_group(world, channels, n, DEV, overlap_sms=overlap)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        def run(owner):
            outs = []
            for r, (peer, st) in enumerate(zip(ranks, streams)):
                b, e = nd_dist.view_shard(nv, r, world)
                with torch.cuda.stream(st):
                    outs.append(peer(f[b:e], pts, proj[b:e], nv, owner=owner))
            torch.cuda.synchronize()
            return outs
        full = [(m.clone(), c.clone(), k.clone()) for m, c, k in run(-1)]
        for peer in ranks:
            peer.mean.fill_(-7.0)
            peer.cov.fill_(-7.0)
        for scene in range(world + 1):
            owner = scene % world
            outs = run(owner)
            for r, o in enumerate(outs):
                if r == owner:
                    assert torch.equal(o[0], full[r][0]) and torch.equal(o[1], full[r][1]) and torch.equal(o[2], full[r][2])
                else:
                    assert o[0] is None and o[1] is None
            if scene == 0:                                   # nobody but the owner was written to
                for r, peer in enumerate(ranks):
                    if r != owner:
                        assert float(peer.mean.max()) == -7.0 and float(peer.cov.min()) == -7.0
        for peer in ranks:
            peer.check()
    finally:
        for peer in ranks:
            peer.close()


@pytest.mark.parametrize('world,overlap', [(2, 0), (8, 0), (2, 3), (8, 2)])
def test_in_process_ranks_without_variance(world, overlap):
    """want_cov=False: the accumulators are [S1 | count], the exchange moves half the bytes and only the mean comes back."""
    nv, grid, channels = 19, (16, 16, 8), 20
    n = int(np.prod(grid))
    f, pts, proj = _scene(nv, grid, (0.4, 0.4, 0.4), channels, 93)
    mean, _, cnt = lifting.lift_mean_var(f, pts, proj)
    ranks = nd_dist.PeerLift.local_group(world, channels, n, DEV, want_cov=False, overlap_sms=overlap)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        torch.cuda.synchronize()
        for epoch in range(2):
            outs = []
            for r, (peer, st) in enumerate(zip(ranks, streams)):
                b, e = nd_dist.view_shard(nv, r, world)
                with torch.cuda.stream(st):
                    outs.append(peer(f[b:e], pts, proj[b:e], nv))
            torch.cuda.synchronize()
        for peer in ranks:
            peer.check()
        for r, (m2, c2, n2) in enumerate(outs):
            assert c2 is None
            assert torch.equal(n2, cnt), f'count of rank {r}'
            assert_close(m2, mean, 1e-4, f'mean of rank {r}')
            assert torch.equal(m2, outs[0][0])
    finally:
        for peer in ranks:
            peer.close()


def test_missing_peer_times_out_instead_of_hanging():
    """A rank whose peer never reaches the exchange step: after the bounded wait the error word is raised on EVERY
    rank, the rows the waiting rank owns are NaN (no stale or partial sums are handed out), the next call fails
    without a synchronisation, and check() fails."""
    ranks = nd_dist.PeerLift.local_group(2, 4, 64, DEV, timeout_ms=100)
    try:
        ranks[0].acc.fill_(1.0)
        ranks[0].mean.fill_(7.0)
        mean, cov, cnt = ranks[0].exchange(3)
        torch.cuda.synchronize()
        b, e = nd_dist.channel_shard(4, 0, 2)
        assert torch.isnan(mean[b:e]).all() and torch.isnan(cov[b:e]).all()     # own slice poisoned
        assert (mean[e:] == 7.0).all()                                          # the peer's slice was never written
        w = 2 * ranks[0]._lib.ND_MAX_PEERS + 1
        assert int(ranks[0].flags[w]) == 1 and int(ranks[1].flags[w]) == 1      # raised in both flag blocks
        with pytest.raises(RuntimeError, match='time-out'):
            ranks[0].exchange(3)                                                # hard error at the next call
        with pytest.raises(RuntimeError, match='peer'):
            ranks[0].check()
    finally:
        for peer in ranks:
            peer.close()
