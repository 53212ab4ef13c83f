"""ctypes binding of oracle/brax_step.c: the scalar C (OpenMP over envs) restatement of
oracle/brax_v1.System.step / System.info.

TEST INFRASTRUCTURE ONLY (see oracle/threefry.py header): the CPU baseline of bench.py (`cpu_baseline`,
`--impl reference`) and a second checker in tests/; never imported by po_brax_b200/.

`attach(system, threads)` makes an existing `brax_v1.System` instance evaluate `step` / `info` in C (same
arguments, same results to rounding; tests/test_oracle_c.py holds float64 to 1e-11 and float32 to the parity
tolerances against the NumPy text). Reference call sites: `self.sys.step(state.qp, action)` at
/root/reference/po_brax/envs/ant_heavenhell.py:108, ant_gather.py:127, ant_tag.py:109; `self.sys.info(qp)` at
ant_heavenhell.py:77, ant_gather.py:95, ant_tag.py:81.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import brax_v1 as bx

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'brax_step.c')
HDR = os.path.join(HERE, 'brax_step_impl.h')
OUT = os.path.join(HERE, '_build', 'libbraxstep.so')
_lib = None
N_CAUSE = 9   # BRAX_NCAUSE
CAUSES = ('ground pen', 'ground nv', 'ground J', 'ground |v_d|', 'wall pen', 'wall nv', 'wall J', 'wall |v_d|',
          'actuator cut-off')


def build(force=False):
    """gcc -O3 -ffp-contract=off -fopenmp (baseline x86-64: the .so travels to the GPU box)."""
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['gcc', '-O3', '-fno-math-errno', '-fno-trapping-math', '-ffp-contract=off', '-fopenmp', '-Wall', '-shared',
           '-fPIC', SRC, '-o', OUT, '-lm']
    subprocess.run(cmd, check=True)
    return OUT


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.brax_step_max_threads.restype = C.c_int
    return _lib


def _desc_type(real):
    p = C.c_void_p
    return type('SysDesc', (C.Structure,), {'_fields_': [
        *[(n, C.c_int) for n in ('nb', 'nj', 'na', 'ncp', 'ncap', 'nbox', 'substeps', 'ground', 'arena')],
        *[(n, real) for n in ('h', 'vel_damp', 'ang_damp', 'baumgarte', 'friction', 'elasticity')],
        ('gravity', real * 3),
        *[(n, p) for n in ('mass', 'inv_inertia', 'active', 'j_parent', 'j_child', 'j_off_p', 'j_off_c', 'j_stiff',
                           'j_sdamp', 'j_adamp', 'j_lstr', 'j_limit', 'j_axis', 'a_joint', 'a_strength', 'cp_body',
                           'cp_end', 'cp_rad', 'cap_body', 'cap_a', 'cap_b', 'cap_rad', 'boxes')]]})


class CBackend:
    """Flat description of a brax_v1.System for the C step; keeps the arrays it points to alive."""

    def __init__(self, system: bx.System, threads: int = 1):
        s = self.system = system
        self.dtype = np.dtype(s.dtype)
        if self.dtype == np.float32:
            real, self.sfx = C.c_float, '_f32'
        elif self.dtype == np.float64:
            real, self.sfx = C.c_double, '_f64'
        else:
            raise TypeError(self.dtype)
        self.threads = int(threads)
        self.nb, self.na = s.num_bodies, len(s.a_joint)
        self.keep = {}

        def arr(name, a, dt):
            a = np.ascontiguousarray(np.asarray(a, dt))
            self.keep[name] = a
            return a.ctypes.data if a.size else None

        d = _desc_type(real)()
        d.nb, d.nj, d.na = s.num_bodies, s.num_joints, len(s.a_joint)
        d.ncp, d.ncap, d.nbox = len(s.cp_body), len(s.cap_body), len(s.boxes)
        d.substeps = s.substeps
        d.ground = -1 if s.ground is None else s.ground
        d.arena = -1 if s.arena is None else s.arena
        d.h, d.vel_damp, d.ang_damp = float(s.h), float(s.vel_damp), float(s.ang_damp)
        d.baumgarte, d.friction, d.elasticity = float(s.baumgarte), float(s.friction), float(s.elasticity)
        d.gravity = (real * 3)(*[float(x) for x in s.gravity])
        ft, it = self.dtype, np.int32
        for name, a, dt in (
                ('mass', s.mass, ft), ('inv_inertia', s.inv_inertia, ft), ('active', s.active, ft),
                ('j_parent', s.j_parent, it), ('j_child', s.j_child, it), ('j_off_p', s.j_off_p, ft),
                ('j_off_c', s.j_off_c, ft), ('j_stiff', s.j_stiff, ft), ('j_sdamp', s.j_sdamp, ft),
                ('j_adamp', s.j_adamp, ft), ('j_lstr', s.j_lstr, ft), ('j_limit', s.j_limit, ft),
                ('j_axis', s.j_axis, ft), ('a_joint', s.a_joint, it), ('a_strength', s.a_strength, ft),
                ('cp_body', s.cp_body, it), ('cp_end', s.cp_end, ft), ('cp_rad', s.cp_rad, ft),
                ('cap_body', s.cap_body, it), ('cap_a', s.cap_a, ft), ('cap_b', s.cap_b, ft),
                ('cap_rad', s.cap_rad, ft), ('boxes', s.boxes, ft)):
            setattr(d, name, arr(name, a, dt))
        self.desc = d
        lib = load()
        self._step = getattr(lib, 'brax_step' + self.sfx)
        self._info = getattr(lib, 'brax_info' + self.sfx)
        for f in (self._step, self._info):
            f.restype = C.c_int

    def _c(self, a):
        return np.ascontiguousarray(a, self.dtype)

    def step(self, qp: bx.QP, act, flip_mask=None, flip_thr=0.0, want_causes=False):
        """System.step(qp, act) -> (qp', Info); the inputs are not modified.

        Two-branch test aid (brax_step_impl.h BranchCtl): with `flip_thr` > 0 the step also counts, per env, the
        evaluations of a discontinuous predicate that were closer than flip_thr to their switching point
        (-> self.n_marginal [N]); `flip_mask` uint32 [N] inverts the k-th such evaluation where bit k is set;
        want_causes -> self.cause_margin [N, 9] (ground pen/nv/J/nd, wall pen/nv/J/nd, actuator cut-off)."""
        pos, rot, vel, ang = (self._c(x).copy() for x in (qp.pos, qp.rot, qp.vel, qp.ang))
        act = self._c(act)
        n = pos.shape[0]
        assert pos.shape == (n, self.nb, 3) and rot.shape == (n, self.nb, 4) and act.shape == (n, self.na)
        cv, ca = np.empty_like(pos), np.empty_like(pos)
        mg = np.empty(n, np.float64) if self.system.track_margin else None
        ctl = flip_thr > 0 or flip_mask is not None
        fm = None if flip_mask is None else np.ascontiguousarray(flip_mask, np.uint32)
        assert fm is None or fm.shape == (n,)
        nm = np.zeros(n, np.int32) if ctl else None
        cause = np.empty((n, N_CAUSE), np.float64) if want_causes else None

        def ptr(a):
            return C.c_void_p(a.ctypes.data if a is not None and a.size else None)
        rc = self._step(C.byref(self.desc), C.c_long(n), *[ptr(x) for x in (pos, rot, vel, ang, act, cv, ca)],
                        ptr(mg), C.c_int(self.threads), ptr(fm), C.c_double(flip_thr), ptr(nm), ptr(cause))
        if rc:
            raise RuntimeError(f'brax_step{self.sfx} failed: {rc}')
        if mg is not None:   # same protocol as the NumPy text: the caller clears system.margin before a step
            self.system.margin = mg if self.system.margin is None else np.minimum(self.system.margin, mg)
        self.n_marginal, self.cause_margin = nm, cause
        return bx.QP(pos, rot, vel, ang), bx.Info(cv, ca)

    def info(self, qp: bx.QP):
        pos, rot, vel, ang = (self._c(x) for x in (qp.pos, qp.rot, qp.vel, qp.ang))
        n = pos.shape[0]
        cv, ca = np.empty_like(pos), np.empty_like(pos)
        rc = self._info(C.byref(self.desc), C.c_long(n), *[C.c_void_p(x.ctypes.data) for x in
                                                           (pos, rot, vel, ang, cv, ca)], C.c_int(self.threads))
        if rc:
            raise RuntimeError(f'brax_info{self.sfx} failed: {rc}')
        return bx.Info(cv, ca)


def attach(system: bx.System, threads: int = 1) -> CBackend:
    """Route `system.step` / `system.info` through the C restatement (instance-level override)."""
    be = CBackend(system, threads)
    system.step = be.step
    system.info = be.info
    system.c_backend = be
    return be
