"""TEST INFRASTRUCTURE -- generates ``tests/golden/*.npz`` by running the UNMODIFIED
reference (``/root/reference``) on seeded synthetic inputs.  Build container only:
``python -m oracle.make_golden``.  The fixtures hold reference OUTPUTS (plus the
few host-RNG draws that are not reproducible from the seed); inputs are rebuilt
from the same seeds by ``oracle/golden_cases.py`` at test time.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import golden_cases as gc  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def _np(t):
    return t.detach().cpu().numpy()


def gen_lift(case):
    ref = ref_loader.load()
    inp = gc.lift_inputs(case)
    nd = ref.nerfdet
    proj = nd.nerfdet._compute_projection(inp['img_meta'], inp['stride'], None)
    pts = nd.get_points(n_voxels=torch.tensor(inp['n_voxels']),
                        voxel_size=torch.tensor(inp['voxel_size']),
                        origin=torch.tensor(inp['img_meta']['lidar2img']['origin']))
    feats = inp['features_sliced']
    volume, valid = nd.backproject(feats, pts, proj, None, inp['voxel_size'])
    # index-revealing planes: value = y*W + x (exact in fp32) -> the reference's own
    # gather tells us the pixel it picked
    nv, _, h, w = feats.shape
    ramp = torch.arange(h * w, dtype=torch.float32).view(1, 1, h, w).expand(nv, 1, h, w)
    pix_vol, _ = nd.backproject(ramp, pts, proj, None, inp['voxel_size'])
    pix = pix_vol.view(nv, -1).long()
    v2 = valid.view(nv, -1)
    pix[~v2] = -1
    # nerfdet.py:171-181, executed verbatim on the reference's tensors
    volume_sum = volume.sum(dim=0)
    cnt = valid.sum(dim=0)
    volume_mean = volume_sum / (cnt + 1e-8)
    volume_mean[:, cnt[0] == 0] = .0
    volume_cov = torch.sum((volume - volume_mean.unsqueeze(0)) ** 2, dim=0) / (cnt + 1e-8)
    volume_cov[:, cnt[0] == 0] = 1e6
    volume_cov = torch.exp(-volume_cov)
    out = dict(projection=_np(proj), points=_np(pts), valid=_np(v2), pix=_np(pix).astype(np.int32),
               volume_mean=_np(volume_mean), volume_cov=_np(volume_cov), count=_np(cnt))
    if case.get('with_depth'):
        vol_d, valid_d = nd.backproject(feats, pts, proj, inp['depth'], inp['voxel_size'])
        out['valid_depth'] = _np(valid_d.view(nv, -1))
        out['volume_depth_sum'] = _np(vol_d.sum(dim=0))
    return out


def gen_lift_grad(case):
    """nerfdet.py:164-181 under the reference's own autograd: d(sum(mean * g_mean) + sum(cov * g_cov)) / d features."""
    ref = ref_loader.load()
    inp = gc.lift_inputs(case)
    nd = ref.nerfdet
    proj = nd.nerfdet._compute_projection(inp['img_meta'], inp['stride'], None)
    pts = nd.get_points(n_voxels=torch.tensor(inp['n_voxels']), voxel_size=torch.tensor(inp['voxel_size']),
                        origin=torch.tensor(inp['img_meta']['lidar2img']['origin']))
    g_mean, g_cov = gc.lift_grad_upstream(case)
    out = {}
    for tag, use_cov in (('', True), ('_mean_only', False)):
        feats = inp['features_sliced'].clone().requires_grad_(True)
        volume, valid = nd.backproject(feats, pts, proj, inp.get('depth'), inp['voxel_size'])
        volume_sum = volume.sum(dim=0)                     # nerfdet.py:171-181, verbatim
        cov_valid = valid.clone().detach()                 # noqa: F841  (kept like the reference)
        valid = valid.sum(dim=0)
        volume_mean = volume_sum / (valid + 1e-8)
        volume_mean[:, valid[0] == 0] = .0
        volume_cov = torch.sum((volume - volume_mean.unsqueeze(0)) ** 2, dim=0) / (valid + 1e-8)
        volume_cov[:, valid[0] == 0] = 1e6
        volume_cov = torch.exp(-volume_cov)
        loss = (volume_mean * g_mean).sum()
        if use_cov:
            loss = loss + (volume_cov * g_cov).sum()
        loss.backward()
        out['g_features' + tag] = _np(feats.grad)
        if use_cov:
            out['volume_mean'], out['volume_cov'], out['count'] = _np(volume_mean), _np(volume_cov), _np(valid)
    return out


def gen_rays(case):
    """N3: the reference's own get_dtu_raydir (taken out of the unmodified data_augment_utils.py, whose module imports mmcv)
    on the pipeline's pixel grid, and OpenCV's imdenormalize arithmetic (what mmcv.imdenormalize calls)."""
    import ast
    import cv2
    src_path = os.path.join(ref_loader.REFERENCE_ROOT, 'mmdet3d/datasets/pipelines/data_augment_utils.py')
    tree = ast.parse(open(src_path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'get_dtu_raydir'][0]
    ns = {'np': np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), src_path, 'exec'), ns)
    inp = gc.rays_inputs(case)
    k = inp['img_meta']['lidar2img']['intrinsic'].copy()                  # multi_view.py:117-118
    ratio = inp['img_meta']['ori_shape'][0] / inp['img_meta']['img_shape'][0]
    k[:2] = k[:2] / ratio
    h, w, m = inp['height'], inp['width'], inp['margin']
    px, py = np.meshgrid(np.arange(m, w - m).astype(np.float32), np.arange(m, h - m).astype(np.float32))  # :124-127
    pixelcoords = np.stack((px, py), axis=-1).astype(np.float32)
    dirs = [np.reshape(ns['get_dtu_raydir'](pixelcoords, k, r).astype(np.float32), (-1, 3)) for r in inp['camrotc2w']]
    mean = np.array(gc.IMG_NORM['mean']).reshape(1, -1).astype(np.float64)   # multi_view.py:28-29 + mmcv.imdenormalize
    std = np.array(gc.IMG_NORM['std']).reshape(1, -1).astype(np.float64)
    den = []
    for img in inp['img_hwc']:
        t = cv2.multiply(img, std)
        cv2.add(t, mean, t)
        cv2.cvtColor(t, cv2.COLOR_RGB2BGR, t)
        den.append((t.astype(np.uint8) / 255.0).transpose(2, 0, 1))        # multi_view.py:107-110, formating.py:88
    return dict(raydirs=np.stack(dirs), denorm=np.stack(den).astype(np.float32))


def gen_extract(case):
    ref = ref_loader.load()
    inp = gc.extract_inputs(case)
    det = ref_loader.build_reference_detector(
        inp['features'], inp['n_voxels'], inp['voxel_size'], inp['aabb'], inp['near_far_range'],
        inp['N_samples'], inp['N_rand'])
    sd = inp['state']
    det.nerf_mlp.load_state_dict({k: v for k, v in sd.items() if not k.startswith('mapping.')})
    det.mapping.load_state_dict({'0.weight': sd['mapping.0.weight'], '0.bias': sd['mapping.0.bias']})
    det.eval()
    # fresh host RNG state exactly like a fresh import of render_ray.py (line 20)
    ref.render_ray.rng = np.random.RandomState(234)
    torch.manual_seed(case['torch_seed'])
    img = torch.zeros((1, inp['features'].shape[0], 3) + tuple(inp['pad_shape']))
    rb = inp['ray_batch']
    with torch.no_grad():
        x, valids, _, rgb_preds, _ = det.extract_feat(img, [inp['img_meta']], 'train', None, rb)
    oc = rgb_preds[0]['outputs_coarse']
    # the draws the reference made (re-derived from identical generator states)
    total = int((rb['gt_depth'].view(-1) > 0).sum())
    sel = np.random.RandomState(234).choice(total, size=(inp['N_rand'],), replace=False)
    torch.manual_seed(case['torch_seed'])
    t_rand = torch.rand(inp['N_rand'], inp['N_samples'])
    return dict(x=_np(x[0]), valids=_np(valids[0]), rgb=_np(oc['rgb']), depth=_np(oc['depth']),
                weights=_np(oc['weights']), mask=_np(oc['mask']), alpha=_np(oc['alpha']),
                z_vals=_np(oc['z_vals']), transparency=_np(oc['transparency']),
                sigma=_np(rgb_preds[0]['sigma']), gt_rgb=_np(rgb_preds[0]['gt_rgb']),
                gt_depth=_np(rgb_preds[0]['gt_depth']), select_inds=sel.astype(np.int64),
                t_rand=_np(t_rand))


def gen_extract_depth(case):
    """The unmodified reference's extract_feat called WITH a depth prior (nerfdet.py:133-139): both backproject calls are
    gated.  Saves what it returns for the voxel side (x, valids) and the rows it hands to query_density (recorded by
    wrapping the bound method of this one instance -- instrumentation, the reference code is untouched)."""
    ref = ref_loader.load()
    inp = gc.extract_depth_inputs(case)
    det = ref_loader.build_reference_detector(
        inp['features'], inp['n_voxels'], inp['voxel_size'], inp['aabb'], inp['near_far_range'],
        inp['N_samples'], inp['N_rand'])
    sd = inp['state']
    det.nerf_mlp.load_state_dict({k: v for k, v in sd.items() if not k.startswith('mapping.')})
    det.mapping.load_state_dict({'0.weight': sd['mapping.0.weight'], '0.bias': sd['mapping.0.bias']})
    det.eval()
    seen = {}
    query = det.nerf_mlp.query_density

    def recording_query(points, features):
        seen['global_volume'] = features.detach().clone()
        return query(points, features)
    det.nerf_mlp.query_density = recording_query
    ref.render_ray.rng = np.random.RandomState(234)
    torch.manual_seed(case['torch_seed'])
    img = torch.zeros((1, inp['features'].shape[0], 3) + tuple(inp['pad_shape']))
    with torch.no_grad():
        x, valids, _, _, _ = det.extract_feat(img, [inp['img_meta']], 'train', inp['depth'].unsqueeze(0), inp['ray_batch'])
    return dict(x=_np(x[0]), valids=_np(valids[0]), global_volume=_np(seen['global_volume']))


def gen_render_det(case):
    ref = ref_loader.load()
    inp = gc.render_inputs(case)
    field = ref.nerf_mlp.VanillaNeRFRadianceField(
        net_depth=4, net_width=256, skip_layer=3, feature_dim=70,
        net_depth_condition=1, net_width_condition=128)
    field.load_state_dict({k: v for k, v in inp['state'].items() if not k.startswith('mapping.')})
    rr = ref.render_ray
    with torch.no_grad():
        ret = rr.render_rays_func(
            inp['ray_o'], inp['ray_d'], None, None, inp['featmaps'], inp['images'],
            inp['aabb'], inp['near_far_range'], inp['N_samples'], inp['ray_o'].shape[0],
            field, inp['img_meta'], ref.projection.Projector(), 'image', 3, False, 0, True)
        # also the intermediates, from the reference's own helpers
        pts, z = rr.sample_along_camera_ray(inp['ray_o'], inp['ray_d'], inp['near_far_range'],
                                            inp['N_samples'], det=True)
        cams = rr._compute_projection(inp['img_meta'])
        imgs = inp['images'].permute(0, 2, 3, 1).unsqueeze(0)
        rgb_feat, mask = ref.projection.Projector().compute(pts, imgs, cams, inp['featmaps'])
        mean, var = rr.compute_mask_points(rgb_feat, mask)
        pix, front = ref.projection.Projector().compute_projections(pts, cams[0])
    oc = ret['outputs_coarse']
    return dict(rgb=_np(oc['rgb']), depth=_np(oc['depth']), weights=_np(oc['weights']),
                mask=_np(oc['mask']), alpha=_np(oc['alpha']), z_vals=_np(oc['z_vals']),
                transparency=_np(oc['transparency']), sigma=_np(ret['sigma']),
                pts=_np(pts), cameras=_np(cams), view_mask=_np(mask[..., 0]).astype(np.uint8),
                mean=_np(mean.squeeze(2)), expvar=_np(var.squeeze(2)),
                pixel_locations=_np(pix), in_front=_np(front))


def gen_render_grad(case):
    """N1: the reference's render_rays_func under its own autograd -- d(sum(rgb * g_rgb) + sum(depth * g_depth)) with
    respect to the mapped feature maps and the field's weights (a selection of them is stored)."""
    ref = ref_loader.load()
    inp = gc.render_inputs(case)
    field = ref.nerf_mlp.VanillaNeRFRadianceField(
        net_depth=4, net_width=256, skip_layer=3, feature_dim=70,
        net_depth_condition=1, net_width_condition=128)
    field.load_state_dict({k: v for k, v in inp['state'].items() if not k.startswith('mapping.')})
    feat = inp['featmaps'].clone().requires_grad_(True)
    ret = ref.render_ray.render_rays_func(
        inp['ray_o'], inp['ray_d'], None, None, feat, inp['images'],
        inp['aabb'], inp['near_far_range'], inp['N_samples'], inp['ray_o'].shape[0],
        field, inp['img_meta'], ref.projection.Projector(), 'image', 3, False, 0, True)
    g_rgb, g_depth = gc.render_grad_upstream(case)
    oc = ret['outputs_coarse']
    ((oc['rgb'] * g_rgb).sum() + (oc['depth'] * g_depth).sum()).backward()
    out = {'g_featmaps': _np(feat.grad), 'rgb': _np(oc['rgb']), 'depth': _np(oc['depth'])}
    params = dict(field.named_parameters())
    for k in gc.RENDER_GRAD_KEYS:
        out['g_' + k] = _np(params[k].grad)
    return out


def gen_mlp(case):
    ref = ref_loader.load()
    inp = gc.mlp_inputs(case)
    field = ref.nerf_mlp.VanillaNeRFRadianceField(
        net_depth=4, net_width=256, skip_layer=3, feature_dim=70,
        net_depth_condition=1, net_width_condition=128)
    field.load_state_dict({k: v for k, v in inp['state'].items() if not k.startswith('mapping.')})
    with torch.no_grad():
        rgb, sigma = field(inp['pts'], inp['ray_d'], inp['feats'])
        dens = field.query_density(inp['pts'].reshape(-1, 3), inp['feats'].reshape(-1, 70))
        enc = field.posi_encoder(inp['pts'])
    return dict(rgb=_np(rgb), sigma=_np(sigma), density=_np(dens), posenc=_np(enc))


def gen_volume_lookup(case):
    ref = ref_loader.load()
    inp = gc.volume_lookup_inputs(case)
    with torch.no_grad():
        feats, masks = ref.render_ray.volume_sampling(inp['pts'], inp['volume'], inp['aabb'])
    return dict(features=_np(feats), inside=_np(masks))


GENERATORS = dict(lift=gen_lift, lift_grad=gen_lift_grad, rays=gen_rays, render_grad=gen_render_grad, extract=gen_extract, extract_depth=gen_extract_depth, render_det=gen_render_det, mlp=gen_mlp,
                  volume_lookup=gen_volume_lookup)


def main():
    if not ref_loader.available():
        raise SystemExit('reference not available: golden vectors can only be made in the build container')
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = sys.argv[1:]                                   # optional case names: regenerate these only
    for name, case in gc.CASES.items():
        if only and name not in only:
            continue
        data = GENERATORS[case['kind']](case)
        path = os.path.join(OUT, f'{name}.npz')
        np.savez_compressed(path, **data)
        print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB  keys={sorted(data)}')


if __name__ == '__main__':
    main()
