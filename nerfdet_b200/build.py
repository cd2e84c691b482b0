"""Builds ``nerfdet_b200/lib/libnerfdet_lift.so`` (the C-ABI library, include/nerfdet_lift.h)
from ``nerfdet_b200/csrc/*.cu`` with nvcc for sm_100a.  In-tree so the ``.so`` travels to
the GPU box; nothing is JIT-compiled at run time."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from typing import List

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
OBJ_DIR = os.path.join(LIB_DIR, 'obj')
LIB_PATH = os.path.join(LIB_DIR, 'libnerfdet_lift.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr',
    '-diag-suppress', '177',
]


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found; libnerfdet_lift.so cannot be built')


def _sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest() -> str:
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.h')):
                h.update(f.encode())
                with open(os.path.join(root, f), 'rb') as fh:
                    h.update(fh.read())
    return h.hexdigest()


def is_current() -> bool:
    stamp = os.path.join(LIB_DIR, 'build.stamp')
    if not (os.path.isfile(LIB_PATH) and os.path.isfile(stamp)):
        return False
    with open(stamp) as fh:
        return fh.read().strip() == _digest()


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and is_current():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = _sources()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}')
        return obj

    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    for f in os.listdir(OBJ_DIR):                            # objects of sources that no longer exist
        if f.endswith('.o') and os.path.join(OBJ_DIR, f) not in objs:
            os.remove(os.path.join(OBJ_DIR, f))
    cmd = [nvcc, '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('link failed')
    with open(os.path.join(LIB_DIR, 'build.stamp'), 'w') as fh:
        fh.write(_digest())
    return LIB_PATH


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
