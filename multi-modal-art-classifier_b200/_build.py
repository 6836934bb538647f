"""Builds ``libagx.so`` (the C-ABI CUDA library, include/agx.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the resulting
``.so`` is git-ignored but travels with the tree to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_PATH = os.path.join(PKG_DIR, 'libagx.so')
STAMP = os.path.join(PKG_DIR, 'libagx.so.stamp')

SOURCES = ['agx_api.cu', 'agx_csr.cu', 'agx_aggregate.cu', 'agx_gemm.cu', 'agx_gemm_tc.cu',
           'agx_nn.cu', 'agx_gat.cu', 'agx_heads_tc.cu', 'agx_layer.cu', 'agx_comm.cu']

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-DAGX_BUILD']


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _source_hash() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(os.path.dirname(PKG_DIR), 'include', 'agx.h'))
    for f in files:
        with open(f, 'rb') as fh:
            h.update(os.path.basename(f).encode())     # not the absolute path: the tree moves
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every ``csrc/*.cu`` and link ``libagx.so``; returns its path."""
    if not force and is_current():
        return LIB_PATH
    # one builder at a time (torchrun starts one process per GPU in the same tree); the others
    # wait for the lock and then find the library current
    import fcntl
    lock = open(os.path.join(PKG_DIR, '.build.lock'), 'w')
    fcntl.flock(lock, fcntl.LOCK_EX)
    try:
        if not force and is_current():
            return LIB_PATH
        return _build_locked(verbose)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _build_locked(verbose: bool) -> str:
    nvcc = _nvcc()
    objs = []
    procs = []
    obj_dir = os.path.join(PKG_DIR, 'build')
    os.makedirs(obj_dir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}')
    tmp = LIB_PATH + f'.tmp{os.getpid()}'
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', tmp, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode())
        raise RuntimeError('linking libagx.so failed')
    os.replace(tmp, LIB_PATH)             # atomic: a concurrent dlopen never sees a partial file
    with open(STAMP, 'w') as fh:
        fh.write(_source_hash())
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
