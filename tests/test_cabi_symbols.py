"""CPU tier: the C-ABI library builds/loads here (no GPU needed) and exports every
function that include/nerfdet_lift.h declares; the ctypes table covers all of them."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    with open(os.path.join(ROOT, 'include', 'nerfdet_lift.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(nd_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ('nd_version', 'nd_last_error_string', 'nd_project_voxels', 'nd_backproject',
                 'nd_lift_mean_var', 'nd_lift_accumulate', 'nd_lift_finalize', 'nd_lift_finalize_peers',
                 'nd_peer_alloc', 'nd_peer_open', 'nd_peer_close', 'nd_peer_free'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from nerfdet_b200 import _lib
    lib = _lib.load()
    assert lib.nd_version() >= 100
    raw = ctypes.CDLL(_lib.library_path())
    for name in _declared():
        assert hasattr(raw, name), f'{name} declared in the header but not exported'
        assert name in _lib.SIGNATURES, f'{name} has no ctypes signature'
    for name in _lib.SIGNATURES:
        assert name in _declared(), f'{name} bound but not declared in the header'


def test_ops_refuse_cpu_tensors():
    import torch
    from nerfdet_b200 import lifting
    pts = lifting.get_points([2, 2, 2], [1., 1., 1.], [0., 0., 0.])
    with pytest.raises(RuntimeError, match='CUDA'):
        lifting.lift_mean_var(torch.zeros(1, 4, 3, 3), pts, torch.zeros(1, 3, 4))


def test_host_geometry_matches_golden():
    """compute_projection / get_points (host code kept from the reference contract)."""
    import numpy as np
    from nerfdet_b200 import lifting
    from oracle import golden_cases as gc
    for name in ('lift_tiny', 'lift_lowres', 'lift_h240_shift'):
        g = gc.load_golden(name)
        inp = gc.lift_inputs(gc.CASES[name])
        proj = lifting.compute_projection(inp['img_meta'], inp['stride'])
        pts = lifting.get_points(inp['n_voxels'], inp['voxel_size'], inp['img_meta']['lidar2img']['origin'])
        assert np.array_equal(proj.numpy(), g['projection'])
        assert np.array_equal(pts.numpy(), g['points'])


def test_peer_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the multi-GPU exchange entry run before any CUDA call: they must answer with status codes
    (and a message) on a machine without a GPU, never crash."""
    from nerfdet_b200 import _lib
    lib = _lib.load()
    null = ctypes.POINTER(ctypes.c_void_p)()
    args = (1, 0, 1, 4, 2, 8, None, None, None, None, None, 0, 0, -1, None)
    assert lib.nd_lift_finalize_peers(null, null, null, null, *args) == 1                 # ND_ERR_BAD_ARG: no tables
    assert b'null' in lib.nd_last_error_string()
    one = (ctypes.c_void_p * 1)(ctypes.c_void_p(256))
    assert lib.nd_lift_finalize_peers(one, one, None, one, 9, 0, 1, 4, 2, 8, None, None, None, None, None, 0, 0, -1, None) == 1
    assert b'world' in lib.nd_last_error_string()                                         # more ranks than ND_MAX_PEERS
    assert lib.nd_lift_finalize_peers(one, one, None, one, 1, 0, 0, 4, 2, 8, None, None, None, None, None, 0, 0, -1, None) == 1
    assert b'epoch' in lib.nd_last_error_string()                                         # epoch 0 = initial flag state
    assert lib.nd_lift_finalize_peers(one, one, None, one, 1, 0, 1, 4, 2, 8, None, None, None, None, None, 0, 0, 3, None) == 1
    assert b'owner' in lib.nd_last_error_string()                                         # owner outside the world
    assert lib.nd_peer_close(None) == 0 and lib.nd_peer_free(None) == 0


@pytest.mark.parametrize('compiler,flags', [('gcc', ['-std=c99', '-x', 'c']), ('g++', ['-std=c++17', '-x', 'c++'])])
def test_header_is_plain_c_and_cxx(tmp_path, compiler, flags):
    """The boundary is a C ABI: include/nerfdet_lift.h compiles on its own as C99 and as C++ (no torch, no CUDA headers),
    and the ctypes mirrors of its structs have the sizes the C compiler gives them."""
    import shutil
    import subprocess
    if shutil.which(compiler) is None:
        pytest.skip(f'{compiler} not installed')
    from nerfdet_b200 import _lib
    src = tmp_path / 'abi.c'
    src.write_text('#include <stdio.h>\n#include "nerfdet_lift.h"\n'
                   'int main(void) { printf("%zu %zu %zu\\n", sizeof(nd_maps), sizeof(nd_lift_options), sizeof(nd_mlp_weights)); return 0; }\n')
    exe = tmp_path / 'abi'
    r = subprocess.run([compiler, '-Wall', '-Wextra', '-pedantic', '-Werror', f'-I{os.path.join(ROOT, "include")}'] + flags +
                       [str(src), '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [ctypes.sizeof(_lib.NdMaps), ctypes.sizeof(_lib.NdLiftOptions), ctypes.sizeof(_lib.NdMlpWeights)]


def test_cxx_host_example_compiles_and_links_against_the_library(tmp_path):
    """examples/host_plan_lift.cpp -- a C++ host on the raw ABI (plan once per scene, one launch per lift) -- builds against
    the header and links against the in-tree library; without a GPU it must fail with a CUDA error message, not crash."""
    import shutil
    import subprocess
    cuda = '/usr/local/cuda'
    if shutil.which('g++') is None or not os.path.isfile(os.path.join(cuda, 'include', 'cuda_runtime.h')):
        pytest.skip('g++ or the CUDA runtime headers are not installed')
    from nerfdet_b200 import _lib
    libdir = os.path.dirname(_lib.library_path())
    exe = tmp_path / 'host_plan_lift'
    r = subprocess.run(['g++', '-std=c++17', '-Wall', '-Wextra', '-Werror', f'-I{os.path.join(ROOT, "include")}',
                        f'-I{cuda}/include', os.path.join(ROOT, 'examples', 'host_plan_lift.cpp'), f'-L{libdir}',
                        '-lnerfdet_lift', f'-L{cuda}/lib64', '-lcudart', f'-Wl,-rpath,{libdir}', f'-Wl,-rpath,{cuda}/lib64',
                        '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch
    if not torch.cuda.is_available():
        run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
        assert run.returncode == 1 and 'cudaMalloc' in run.stderr and 'ABI version' in run.stdout
