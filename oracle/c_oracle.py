"""TEST INFRASTRUCTURE -- ctypes binding of the plain-C lift oracle (oracle/lift_ref.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liblift_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, 'lift_ref.c')
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def project(points: np.ndarray, proj: np.ndarray, height: int, width: int, want_q: bool = False):
    """points [3, N] f32, proj [nv, 3, 4] f32 -> pix int32 [nv, N] (-1 invalid) [, q [nv,3,N]]."""
    points = np.ascontiguousarray(points.reshape(3, -1), dtype=np.float32)
    proj = np.ascontiguousarray(proj, dtype=np.float32)
    nv, n = proj.shape[0], points.shape[1]
    pix = np.empty((nv, n), dtype=np.int32)
    q = np.empty((nv, 3, n), dtype=np.float32) if want_q else None
    lib().nd_oracle_project(_p(points, ctypes.c_float), _p(proj, ctypes.c_float), ctypes.c_int(nv),
                            ctypes.c_int64(n), ctypes.c_int(height), ctypes.c_int(width),
                            _p(pix, ctypes.c_int32),
                            _p(q, ctypes.c_float) if want_q else None)
    return (pix, q) if want_q else pix


def lift(features: np.ndarray, points: np.ndarray, proj: np.ndarray):
    """features: any-strided f32 view [nv, C, H, W]; returns mean [C,N], cov [C,N], count [N]."""
    assert features.dtype == np.float32
    points = np.ascontiguousarray(points.reshape(3, -1), dtype=np.float32)
    proj = np.ascontiguousarray(proj, dtype=np.float32)
    nv, c, h, w = features.shape
    n = points.shape[1]
    sv, sc, sy, sx = (s // 4 for s in features.strides)
    mean = np.empty((c, n), dtype=np.float32)
    cov = np.empty((c, n), dtype=np.float32)
    cnt = np.empty((n,), dtype=np.int64)
    lib().nd_oracle_lift(ctypes.c_void_p(features.ctypes.data), ctypes.c_int64(sv), ctypes.c_int64(sc),
                         ctypes.c_int64(sy), ctypes.c_int64(sx), ctypes.c_int(nv), ctypes.c_int(c),
                         ctypes.c_int(h), ctypes.c_int(w), _p(points, ctypes.c_float),
                         _p(proj, ctypes.c_float), ctypes.c_int64(n), _p(mean, ctypes.c_float),
                         _p(cov, ctypes.c_float), _p(cnt, ctypes.c_int64))
    return mean, cov, cnt
