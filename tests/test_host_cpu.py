"""CPU tier for the host-side mirror of the reference interface: what the reference keeps on the host
(camera packing, ray selection, module / state_dict layout) must match fixtures made by the reference, and
every op must refuse to run without a CUDA device (no fallback)."""
import inspect

import numpy as np
import pytest
import torch

from nerfdet_b200 import lifting, live, nerf_mlp, ops, projection, render
from nerfdet_b200.synthetic import make_mlp_state
from oracle import golden_cases as gc


def test_camera_packing_matches_reference():
    inp = gc.render_inputs(gc.CASES['render_det'])
    g = gc.load_golden('render_det')
    cams = render._compute_projection(inp['img_meta'])
    assert cams.shape == (1, 6, 34)
    assert np.array_equal(cams.numpy(), g['cameras'])


def test_module_tree_has_reference_state_dict_keys():
    field = nerf_mlp.VanillaNeRFRadianceField(net_depth=4, net_width=256, skip_layer=3, feature_dim=70,
                                              net_depth_condition=1, net_width_condition=128)
    ref_state = {k: v for k, v in make_mlp_state(1).items() if not k.startswith('mapping.')}
    own = field.state_dict()
    assert sorted(own.keys()) == sorted(ref_state.keys())
    for k, v in ref_state.items():
        assert tuple(own[k].shape) == tuple(v.shape), k
    field.load_state_dict(ref_state)                     # strict
    # reference initialisation: zero biases (nerf_mlp.py:60-78)
    fresh = nerf_mlp.VanillaNeRFRadianceField(4, 256, 3, 70)
    assert all(float(p.detach().abs().max()) == 0.0 for n, p in fresh.named_parameters() if n.endswith('bias'))
    # deeper trunk with a mid-network skip (nerf_mlp.py:80-90): layer after the skip takes width + input
    deep = nerf_mlp.VanillaNeRFRadianceField(net_depth=8, net_width=256, skip_layer=4, feature_dim=0)
    assert deep.mlp.base.hidden_layers[5].in_features == 256 + 63
    assert deep.mlp.sigma_layer.output_layer.in_features == 256


def test_signatures_mirror_the_reference():
    """Argument names / order of the reference callables."""
    assert list(inspect.signature(lifting.backproject).parameters) == ['features', 'points', 'projection', 'depth', 'voxel_size']
    assert list(inspect.signature(lifting.get_points).parameters) == ['n_voxels', 'voxel_size', 'origin']
    want = ['ray_batch', 'mean_volume', 'cov_volume', 'features_2D', 'img', 'aabb', 'near_far_range', 'N_samples',
            'N_rand', 'nerf_mlp', 'img_meta', 'projector', 'mode', 'nerf_sample_view', 'inv_uniform', 'N_importance',
            'det', 'is_train', 'white_bkgd', 'render_testing']
    assert list(inspect.signature(render.render_rays).parameters) == want
    assert list(inspect.signature(projection.Projector.compute).parameters) == ['self', 'xyz', 'train_imgs', 'train_cameras',
                                                                               'featmaps', 'grid_sample']
    assert list(inspect.signature(nerf_mlp.VanillaNeRFRadianceField.forward).parameters) == ['self', 'x', 'condition', 'features']
    assert render.rng.randint(1 << 30) == np.random.RandomState(234).randint(1 << 30)


def test_training_ray_selection_matches_reference():
    """R1: flatten, drop gt_depth <= 0, numpy RandomState(234).choice -- checked up to the first CUDA call."""
    case = gc.CASES['extract_small']
    g = gc.load_golden('extract_small')
    inp = gc.extract_inputs(case)
    seen = {}

    def stop(ray_o, ray_d, *args, **kw):
        seen.update(ray_o=ray_o, ray_d=ray_d, gt_rgb=kw['gt_rgb'], gt_depth=kw['gt_depth'])
        raise StopIteration

    render.rng = np.random.RandomState(234)
    orig = render.render_rays_func
    render.render_rays_func = stop
    try:
        with pytest.raises(StopIteration):
            render.render_rays(inp['ray_batch'], None, None, torch.zeros(5, 32, 14, 20), torch.zeros(5, 3, 60, 80), None,
                               inp['near_far_range'], inp['N_samples'], inp['N_rand'], None, inp['img_meta'], None, 'image')
    finally:
        render.render_rays_func = orig
    rb = inp['ray_batch']
    keep = rb['gt_depth'].view(-1) > 0
    sel = g['select_inds']
    assert torch.equal(seen['ray_o'], rb['ray_o'].view(-1, 3)[keep][sel].float())
    assert np.allclose(seen['gt_rgb'].numpy(), g['gt_rgb'])
    assert np.allclose(seen['gt_depth'].numpy(), g['gt_depth'])


def test_new_ops_refuse_cpu_tensors():
    z = torch.zeros
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.sample_rays(z(4, 3), z(4, 3), 0.2, 8.0, 8, None)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.render_gather_stats(z(4, 3), z(2, 34), z(2, 3, 8, 8), z(2, 4, 4, 4), False, False)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.composite(z(2, 4, 3), z(2, 4), z(2, 4), None, z(2), False)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.volume_sample(z(2, 3, 3, 3), z(5, 3), [0., 0., 0.], [1., 1., 1.])
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.live_stats(z(2, 32, 4, 4), z(2, 3, 16, 16), z(3, 8), z(2, 3, 4), z(2, 3, 4), z(32), False)
    field = nerf_mlp.VanillaNeRFRadianceField(4, 256, 3, 70)
    with pytest.raises(RuntimeError, match='CUDA'):
        field.query_density(z(4, 3), z(4, 70))


def test_mlp_packed_size_and_unsupported_architectures():
    import ctypes
    from nerfdet_b200 import _lib
    lib = _lib.load()
    arch = ops.mlp_arch({}, [4, 256, 3, 70, 128, 10, 4])
    # 4 hidden layers (136 / 256 / 256 / 256 rows) + heads with the [h, in] concatenation (392 rows)
    floats = (136 * 256 + 256) + 3 * (256 * 256 + 256) + (392 + 4) + (392 * 256 + 256) + (284 * 128 + 128) + (3 * 128 + 4)
    assert lib.nd_mlp_packed_bytes(ctypes.byref(arch)) == 4 * floats
    assert lib.nd_mlp_packed_bytes(ctypes.byref(ops.mlp_arch({}, [4, 128, 3, 70, 128, 10, 4]))) == 0
    with pytest.raises(NotImplementedError):
        nerf_mlp.VanillaNeRFRadianceField(4, 256, 3, 70, net_depth_condition=2)


def test_reciprocal_quotient_is_correctly_rounded():
    """The lift epilogue (csrc/lift_quads.cu) computes the mean S1 / count as q0 = S1 * rc,
    q = fma(fma(-q0, count, S1), rc, q0) with rc = RN(1 / count).  For the integer divisors 1..254 (uint8 view counts)
    that is the correctly rounded quotient, i.e. bit-equal to the reference's IEEE divide (nerfdet.py:175).  The fma is
    emulated exactly in float64 (a 24-bit by 8-bit product is exact there)."""
    import numpy as np
    rs = np.random.RandomState(0)
    a = (rs.randn(1 << 18) * np.logspace(-6, 6, 1 << 18)).astype(np.float32)
    for n in range(1, 255):
        cf = np.float32(n)
        rc = np.float32(1.0) / cf
        q0 = (a * rc).astype(np.float32)
        r = (a.astype(np.float64) - q0.astype(np.float64) * float(n)).astype(np.float32)
        q = (q0.astype(np.float64) + r.astype(np.float64) * float(rc)).astype(np.float32)
        assert np.array_equal(q, (a / cf).astype(np.float32)), n


def test_ssim_restatement_against_the_definition():
    """oracle/metrics_oracle.ssim (scipy uniform filter, the way scikit-image 0.18.1 evaluates it) against the definition
    written out per pixel: 7 x 7 window means, sample covariance, the map cropped by 3 pixels, mean over channels."""
    from oracle import metrics_oracle as mt
    rs = np.random.RandomState(4)
    a = rs.uniform(0, 1, (15, 17, 3)).astype(np.float32)
    b = np.clip(a + 0.1 * rs.standard_normal(a.shape), 0, 1)
    c1, c2, vals = (0.01 * 2.0) ** 2, (0.03 * 2.0) ** 2, []
    for ch in range(3):
        x, y = a[..., ch].astype(np.float64), b[..., ch].astype(np.float64)
        for i in range(3, 15 - 3):
            for j in range(3, 17 - 3):
                wx, wy = x[i - 3:i + 4, j - 3:j + 4], y[i - 3:i + 4, j - 3:j + 4]
                ux, uy = wx.mean(), wy.mean()
                vx, vy = wx.var(ddof=1), wy.var(ddof=1)
                vxy = ((wx - ux) * (wy - uy)).sum() / 48.0
                vals.append((2 * ux * uy + c1) * (2 * vxy + c2) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2)))
    assert abs(mt.ssim(a, b) - float(np.mean(vals))) < 1e-12
    assert abs(mt.ssim(a, a.astype(np.float64)) - 1.0) < 1e-12


def _run_bench(argv, env_extra=None, timeout=300):
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(root, 'bench.py')] + argv, cwd=root, env=env, timeout=timeout,
                          capture_output=True, text=True)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` needs no GPU: ONE JSON line on rank 0 with the arm's own cpu_baseline and an e2e that
    repeats the line's value with zero copy bytes; the other ranks of a torchrun launch exit 0 without work."""
    import json
    r = _run_bench(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '0'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'voxel_view_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 1 and d['gpu_launches'] == 0
    assert d['value'] > 0 and abs(d['value'] - 50 * 25600 / (d['ms_per_step'] * 1e-3)) <= 1e-6 * d['value']
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] and d['cpu_baseline']['sample']
    assert d['e2e'] == {'value': d['value'], 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['config']['views'] == 50 and d['config']['n_voxels'] == [40, 40, 16]
    r1 = _run_bench(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                    {'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only box check')
def test_bench_product_arm_fails_loudly_without_a_gpu():
    r = _run_bench(['--steps', '1', '--warmup', '0'], timeout=120)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_bench_algorithmic_bytes_and_clock_parsing():
    """The roofline numerator is SURVEY.md section 8d's formula (294.3 MB per scene at configs[1], 535.9 MB at 100 views,
    662.7 MB at 80x80x32) and the clock sampler keeps only samples taken under load."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_under_test', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.algorithmic_bytes(50, 256, 59, 80, 25600) == 50 * 256 * 59 * 80 * 4 + 2 * 256 * 25600 * 4 + 25600 * 8 + 50 * 48
    assert round(bench.algorithmic_bytes(50, 256, 59, 80, 25600) / 1e6, 1) == 294.3
    assert abs(bench.algorithmic_bytes(100, 256, 59, 80, 25600) / 1e6 - 535.9) < 0.1      # SURVEY truncates 535.97
    assert abs(bench.algorithmic_bytes(50, 256, 59, 80, 204800) / 1e6 - 662.7) < 0.1
    s = bench.ClockSampler(0)
    s.proc = type('P', (), {'terminate': lambda self: None, 'wait': lambda self, timeout=None: 0, 'kill': lambda self: None})()
    s.lines = ['345, 1965, 0, Not Active, Not Active, Not Active, Not Active',
               '1965, 1965, 100, Not Active, Not Active, Not Active, Not Active',
               '1785, 1965, 100, Not Active, Not Active, Not Active, Active', 'garbage']
    c = s.stop()
    assert c['sm_mhz'] == 1875.0 and c['sm_max_mhz'] == 1965.0 and c['reasons'] == ['sw_power_cap']
    assert c['samples'] == 4 and c['samples_under_load'] == 2
