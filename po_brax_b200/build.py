"""Builds po_brax_b200/libpobrax.so (sm_100a only) with nvcc, in-tree.

    python -m po_brax_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libpobrax.so')
SOURCES = ['kernels.cu', 'api.cu']
HEADERS = ['ant_physics.cuh', 'dev_const.h', 'threefry.cuh', 'vec.cuh', os.path.join('..', '..', 'include', 'pobrax.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, out=None, extra=()):
    """out / extra: tuning builds (another output path, extra nvcc flags such as -DPOBRAX_WALL_WARPS_PER_SMSP=5)."""
    if out is None and not force and not is_stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + (['-Xptxas', '-v'] if verbose else []) + ['-o', out or LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return out or LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
