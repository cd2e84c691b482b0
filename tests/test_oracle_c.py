"""The plain-C lift oracle (oracle/lift_ref.c) against the reference-generated
fixtures: pixel indices / validity / counts bit-exact, statistics to fp32 round-off."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import golden_cases as gc
from oracle import lift_oracle as lo

CASES = [k for k, v in gc.CASES.items() if v['kind'] == 'lift']


@pytest.mark.parametrize('name', CASES)
def test_c_oracle_matches_reference(name):
    g = gc.load_golden(name)
    inp = gc.lift_inputs(gc.CASES[name])
    f = inp['features_sliced']
    pix, q = c_oracle.project(g['points'], g['projection'], f.shape[2], f.shape[3], want_q=True)
    assert np.array_equal(pix, g['pix'].astype(np.int32))
    # the FMA chain is what torch.bmm computes on this host
    pts = torch.from_numpy(g['points'])
    _, _, _, q_t = lo.project_voxels(pts, torch.from_numpy(g['projection']), f.shape[2], f.shape[3])
    assert np.array_equal(q, q_t.numpy())
    mean, cov, cnt = c_oracle.lift(f.numpy(), g['points'], g['projection'])
    assert np.array_equal(cnt, g['count'].reshape(-1))
    c = f.shape[1]
    np.testing.assert_allclose(mean, g['volume_mean'].reshape(c, -1), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(cov, g['volume_cov'].reshape(c, -1), rtol=2e-5, atol=1e-6)


def test_c_oracle_strided_view():
    """Non-contiguous feature slice (the reference passes feature[:, :, :h, :w])."""
    inp = gc.lift_inputs(gc.CASES['lift_tiny'])
    g = gc.load_golden('lift_tiny')
    full = inp['features'].numpy()
    view = full[:, :, :14, :20]
    assert not view.flags['C_CONTIGUOUS']
    m1, c1, n1 = c_oracle.lift(view, g['points'], g['projection'])
    m2, c2, n2 = c_oracle.lift(np.ascontiguousarray(view), g['points'], g['projection'])
    assert np.array_equal(m1, m2) and np.array_equal(c1, c2) and np.array_equal(n1, n2)
