"""Host (g++) build of the device physics source, for the CPU suite (see emu.cpp). TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, 'emu.cpp')
OUT = os.path.join(HERE, '_build', 'libpobrax_emu.so')
DEPS = [SRC, os.path.join(HERE, 'shim.h')] + [os.path.join(ROOT, 'po_brax_b200', 'csrc', f) for f in
                ('api.cu', 'ant_physics.cuh', 'vec.cuh', 'dev_const.h', 'threefry.cuh')] + \
       [os.path.join(ROOT, 'include', 'pobrax.h')]
CUDA = os.environ.get('CUDA_HOME', '/usr/local/cuda')
_lib = None


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['g++', '-std=c++17', '-O2', '-ffp-contract=off', '-fPIC', '-shared', '-Wall', '-Wno-unused-function',
           '-Wno-unknown-pragmas', '-I' + os.path.join(CUDA, 'include'), '-x', 'c++', SRC, '-o', OUT,
           '-L' + os.path.join(CUDA, 'lib64'), '-lcudart_static', '-ldl', '-lrt', '-lpthread', '-Wl,-Bsymbolic']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('host emulator build failed:\n' + ' '.join(cmd) + '\n' + r.stdout + r.stderr)
    return OUT


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emu_create.restype = C.c_int
        _lib.emu_step.restype = C.c_int
        _lib.emu_num_bodies.restype = C.c_int
        _lib.emu_tip_masks.restype = C.c_int
        _lib.emu_walls.restype = C.c_int
        _lib.emu_destroy.restype = None
        _lib.pobrax_last_error.restype = C.c_char_p
    return _lib
