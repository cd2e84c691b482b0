"""The CPU oracle (oracle/*.py) against fixtures produced by the unmodified
reference (tests/golden/*.npz, made by oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import golden_cases as gc
from oracle import lift_oracle as lo
from oracle import mlp_oracle as mo
from oracle import render_oracle as ro

LIFT_CASES = [k for k, v in gc.CASES.items() if v['kind'] == 'lift']


def _close(a, b, rtol=1e-5, atol_scale=1e-6, name=''):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    atol = atol_scale * max(1.0, float(np.abs(b).max()))
    bad = np.abs(a - b) > rtol * np.abs(b) + atol
    assert not bad.any(), f'{name}: {bad.sum()} / {bad.size} differ, max abs {np.abs(a - b).max():.3e}'


@pytest.mark.parametrize('name', LIFT_CASES)
def test_lift_oracle_matches_reference(name):
    case = gc.CASES[name]
    g = gc.load_golden(name)
    inp = gc.lift_inputs(case)
    proj = lo.compute_projection(inp['img_meta'], inp['stride'])
    pts = lo.get_points(inp['n_voxels'], inp['voxel_size'], inp['img_meta']['lidar2img']['origin'])
    assert np.array_equal(proj.numpy(), g['projection'])
    assert np.array_equal(pts.numpy(), g['points'])
    f = inp['features_sliced']
    x, y, valid, _ = lo.project_voxels(pts, proj, f.shape[2], f.shape[3])
    assert np.array_equal(valid.numpy(), g['valid'])
    pix = (y * f.shape[3] + x)
    pix[~valid] = -1
    assert np.array_equal(pix.numpy(), g['pix'].astype(np.int64))
    assert 0.02 < valid.float().mean() < 0.98          # the case is not degenerate
    mean, cov, cnt = lo.lift_mean_var(f, pts, proj)
    assert np.array_equal(cnt.numpy(), g['count'])
    _close(mean, g['volume_mean'], name='mean')
    _close(cov, g['volume_cov'], name='cov')
    if case.get('with_depth'):
        vol, vd = lo.backproject(f, pts, proj, inp['depth'], inp['voxel_size'])
        assert np.array_equal(vd.view(vd.shape[0], -1).numpy(), g['valid_depth'])
        _close(vol.sum(dim=0), g['volume_depth_sum'], name='depth volume')
        assert vd.sum() < valid.sum()


@pytest.mark.parametrize('name', [k for k, v in gc.CASES.items() if v['kind'] == 'lift_grad'])
def test_lift_oracle_autograd_matches_reference_autograd(name):
    """Row N1: the oracle's lift is differentiable torch code; its gradient with respect to the features equals
    the one the unmodified reference's autograd produced (fixtures of oracle/make_golden.py:gen_lift_grad)."""
    case = gc.CASES[name]
    g = gc.load_golden(name)
    inp = gc.lift_inputs(case)
    proj = lo.compute_projection(inp['img_meta'], inp['stride'])
    pts = lo.get_points(inp['n_voxels'], inp['voxel_size'], inp['img_meta']['lidar2img']['origin'])
    g_mean, g_cov = gc.lift_grad_upstream(case)
    for tag, use_cov in (('', True), ('_mean_only', False)):
        f = inp['features_sliced'].clone().requires_grad_(True)
        volume, valid = lo.backproject(f, pts, proj, inp.get('depth'), inp['voxel_size'])
        mean, cov, cnt = lo.mean_var(volume, valid)
        loss = (mean * g_mean).sum() + ((cov * g_cov).sum() if use_cov else 0.0)
        loss.backward()
        assert np.array_equal(cnt.numpy(), g['count'])
        _close(f.grad, g['g_features' + tag], rtol=1e-4, atol_scale=1e-6, name='g_features' + tag)
        assert float(f.grad.abs().max()) > 0


def test_rays_oracle_matches_reference_and_opencv():
    """Row N3: ray directions against the reference's own get_dtu_raydir, de-normalised images against OpenCV's
    arithmetic (what mmcv.imdenormalize calls) -- both bit for bit."""
    from oracle import rays_oracle as rys
    case = gc.CASES['rays_small']
    g = gc.load_golden('rays_small')
    inp = gc.rays_inputs(case)
    k = inp['img_meta']['lidar2img']['intrinsic'].copy()
    k[:2] = k[:2] / (inp['img_meta']['ori_shape'][0] / inp['img_meta']['img_shape'][0])
    for t, rot in enumerate(inp['camrotc2w']):
        d = rys.raydirs(k, rot, inp['height'], inp['width'], inp['margin'])
        assert np.array_equal(d, g['raydirs'][t])
    for i, img in enumerate(inp['img_hwc']):
        den = rys.denorm(img, gc.IMG_NORM['mean'], gc.IMG_NORM['std'], True).transpose(2, 0, 1).astype(np.float32)
        assert np.array_equal(den, g['denorm'][i])
    assert 0.0 <= g['denorm'].min() and g['denorm'].max() <= 1.0 and g['denorm'].std() > 0.1


def test_render_oracle_autograd_matches_reference_autograd():
    """Row N1, render branch: torch autograd through the oracle (grid_sample, masked statistics, MLP, compositing) gives
    the gradients the unmodified reference's render_rays_func gave (fixture of oracle/make_golden.py:gen_render_grad)."""
    case = gc.CASES['render_grad']
    g = gc.load_golden('render_grad')
    inp = gc.render_inputs(case)
    so = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in inp['state'].items()}
    fo = inp['featmaps'].clone().requires_grad_(True)
    out = ro.render_image_mode(inp['ray_o'], inp['ray_d'], fo, inp['images'], inp['near_far_range'], case['N_samples'],
                               mo.FieldOracle(so), inp['img_meta'], det=True)['outputs_coarse']
    _close(out['rgb'].detach(), g['rgb'], name='rgb')
    _close(out['depth'].detach(), g['depth'], name='depth')
    g_rgb, g_depth = gc.render_grad_upstream(case)
    ((out['rgb'] * g_rgb).sum() + (out['depth'] * g_depth).sum()).backward()
    _close(fo.grad, g['g_featmaps'], rtol=1e-3, atol_scale=1e-5, name='g_featmaps')
    for k in gc.RENDER_GRAD_KEYS:
        _close(so[k].grad, g['g_' + k], rtol=1e-3, atol_scale=1e-5, name=k)
    assert float(np.abs(g['g_featmaps']).max()) > 0


def test_mlp_oracle_matches_reference():
    g = gc.load_golden('mlp_small')
    inp = gc.mlp_inputs(gc.CASES['mlp_small'])
    field = mo.FieldOracle(inp['state'])
    assert np.array_equal(mo.sinusoidal_encode(inp['pts'], 10).numpy(), g['posenc'])
    rgb, sigma = field(inp['pts'], inp['ray_d'], inp['feats'])
    _close(rgb, g['rgb'], name='rgb')
    _close(sigma, g['sigma'], name='sigma')
    dens = field.query_density(inp['pts'].reshape(-1, 3), inp['feats'].reshape(-1, 70))
    _close(dens, g['density'], name='density')
    assert (g['sigma'] > 0).mean() > 0.1


def test_render_oracle_matches_reference():
    case = gc.CASES['render_det']
    g = gc.load_golden('render_det')
    inp = gc.render_inputs(case)
    cams = ro.pack_cameras(inp['img_meta'])
    assert np.array_equal(cams.numpy(), g['cameras'])
    pts, z = ro.sample_along_rays(inp['ray_o'], inp['ray_d'], *inp['near_far_range'],
                                  inp['N_samples'], det=True)
    assert np.array_equal(z.numpy(), g['z_vals'])
    assert np.array_equal(pts.numpy(), g['pts'])
    pix, front = ro.project_samples(pts, cams[0])
    assert np.array_equal(front.numpy(), g['in_front'])
    assert np.array_equal(pix.numpy(), g['pixel_locations'])
    feat, mask = ro.gather_views(pts, inp['images'], cams[0], inp['featmaps'])
    assert np.array_equal(mask[..., 0].numpy().astype(np.uint8), g['view_mask'])
    mean, var = ro.view_statistics(feat, mask)
    _close(mean.squeeze(2), g['mean'], name='mean35')
    _close(var.squeeze(2), g['expvar'], name='expvar35')
    out = ro.render_image_mode(inp['ray_o'], inp['ray_d'], inp['featmaps'], inp['images'],
                               inp['near_far_range'], inp['N_samples'],
                               mo.FieldOracle(inp['state']), inp['img_meta'], det=True)
    oc = out['outputs_coarse']
    assert np.array_equal(oc['mask'].numpy(), g['mask'])
    for k in ('rgb', 'depth', 'weights', 'alpha', 'transparency'):
        _close(oc[k], g[k], rtol=2e-5, name=k)
    _close(out['sigma'], g['sigma'], name='sigma')
    assert 0.05 < g['view_mask'].mean() < 0.95


def test_volume_lookup_oracle_matches_reference():
    g = gc.load_golden('volume_lookup')
    inp = gc.volume_lookup_inputs(gc.CASES['volume_lookup'])
    feats, inside = ro.volume_lookup(inp['pts'], inp['volume'], inp['aabb'])
    assert np.array_equal(inside.numpy(), g['inside'])
    _close(feats, g['features'], name='trilinear')
    assert 0.1 < g['inside'].mean() < 0.9


def test_extract_feat_oracle_matches_reference():
    """Whole voxel side + render branch of nerfdet.extract_feat, train mode."""
    case = gc.CASES['extract_small']
    g = gc.load_golden('extract_small')
    inp = gc.extract_inputs(case)
    sd = inp['state']
    field = mo.FieldOracle(sd)
    meta = inp['img_meta']
    h, w = meta['img_shape'][0] // 4, meta['img_shape'][1] // 4
    fs = inp['features'][:, :, :h, :w]
    res = lo.extract_lift(fs, meta, inp['n_voxels'], inp['voxel_size'],
                          inp['ray_batch']['denorm_images'], sd['mapping.0.weight'],
                          sd['mapping.0.bias'], field)
    assert np.array_equal(res['count'].numpy(), g['valids'])
    _close(res['x_scene'], g['x'], rtol=1e-4, atol_scale=1e-5, name='x_scene')
    assert (g['valids'] == 0).any() and (g['valids'] > 0).any()
    # render branch with the reference's host draws
    ray_o, ray_d, gt_rgb, gt_depth, sel = ro.select_training_rays(
        inp['ray_batch'], inp['N_rand'], np.random.RandomState(234))
    assert np.array_equal(sel, g['select_inds'])
    imgs = inp['ray_batch']['denorm_images'][0]
    out = ro.render_image_mode(ray_o, ray_d, res['feature_2d'], imgs, inp['near_far_range'],
                               inp['N_samples'], field, meta, det=False,
                               t_rand=torch.from_numpy(g['t_rand']))
    oc = out['outputs_coarse']
    assert np.array_equal(oc['z_vals'].numpy(), g['z_vals'])
    assert np.array_equal(oc['mask'].numpy(), g['mask'])
    for k in ('rgb', 'depth', 'weights', 'alpha', 'transparency'):
        _close(oc[k], g[k], rtol=1e-4, atol_scale=1e-5, name=k)
    _close(gt_rgb, g['gt_rgb'], name='gt_rgb')
    _close(gt_depth, g['gt_depth'], name='gt_depth')


def test_extract_feat_oracle_with_depth_prior_matches_reference():
    """extract_feat(depth=...): backproject's gate (row B4) on the 256-channel volume and on the RGB volume, hence on the
    live statistics and the density.  The fixture holds what the unmodified reference returned (x, valids) and the rows it
    handed to query_density."""
    case = gc.CASES['extract_depth']
    g = gc.load_golden('extract_depth')
    inp = gc.extract_depth_inputs(case)
    sd = inp['state']
    meta = inp['img_meta']
    h, w = meta['img_shape'][0] // 4, meta['img_shape'][1] // 4
    args = (inp['features'][:, :, :h, :w], meta, inp['n_voxels'], inp['voxel_size'], inp['ray_batch']['denorm_images'],
            sd['mapping.0.weight'], sd['mapping.0.bias'], mo.FieldOracle(sd))
    res = lo.extract_lift(*args, depth=inp['depth'])
    assert np.array_equal(res['count'].numpy(), g['valids'])
    obs = g['valids'].reshape(-1) > 0
    assert obs.any() and (~obs).any()
    _close(res['global_volume'][torch.from_numpy(obs)], g['global_volume'][obs], rtol=1e-4, atol_scale=1e-5, name='global_volume')
    _close(res['x_scene'], g['x'], rtol=1e-4, atol_scale=1e-5, name='x_scene')
    ungated = lo.extract_lift(*args)                      # the gate must matter in this fixture
    assert int(ungated['count'].sum()) > 2 * int(g['valids'].sum())


def test_bf16_mlp_oracle_within_bf16_tolerance_of_reference():
    """oracle/mlp_oracle.py:FieldOracleBf16 (the operand roundings of the tensor-core kernel) stays within
    BASELINE.json's bf16 tolerance (1e-2, normalised by the tensor's max) of the reference fixture."""
    import numpy as np
    from oracle import golden_cases as gc
    from oracle import mlp_oracle as mo
    g = gc.load_golden('mlp_small')
    inp = gc.mlp_inputs(gc.CASES['mlp_small'])
    rgb, sigma = mo.FieldOracleBf16(inp['state'])(inp['pts'], inp['ray_d'], inp['feats'])
    for name, a, b in (('rgb', rgb.numpy(), g['rgb']), ('sigma', sigma.numpy(), g['sigma'])):
        err = np.abs(a - b).max()
        assert 0 < err <= 1e-2 * np.abs(b).max(), (name, err)
