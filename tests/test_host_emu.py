"""The DEVICE physics source (po_brax_b200/csrc/ant_physics.cuh: advance2 / substep2 / contacts2, the packed
float32x2 arithmetic, wall candidate tables, out-of-line contact groups) compiled for the host by g++ and run
against the oracle -- no GPU (tests/host_emu/emu.cpp explains how the quad shuffles, the texture lookup and the
PTX are stood in for). It checks the TEXT of what the step kernels run; the parity tests proper are the `-m gpu`
ones, which run the real kernels through the C ABI.

Same gates as the GPU teacher-forced tests (tests/_parity.py): pos / rot 1e-6 + 2e-6|x|, vel / ang / contact impulses
3e-4 on every env, where an env with rounding-ambiguous contact / actuator decisions in that step may instead meet
them against the oracle with those decisions taken the other way (two-branch check)."""
import ctypes as C

import numpy as np
import pytest

from oracle import envs as oenvs
from po_brax_b200 import _lib
from tests import _parity as P
from tests import host_emu

KINDS = {'ant': _lib.ANT, 'ant_heavenhell': _lib.ANT_HEAVENHELL, 'ant_tag': _lib.ANT_TAG, 'ant_gather': _lib.ANT_GATHER}


class Emu:
    def __init__(self, kind, **params):
        self.lib = host_emu.load()
        p = _lib.PobraxParams()
        assert _lib.load().pobrax_default_params(KINDS[kind], C.byref(p)) == 0   # host code of the product library
        p.num_envs = 1
        for k, v in params.items():
            setattr(p, k, v)
        self.h = C.c_void_p()
        rc = self.lib.emu_create(C.byref(p), C.byref(self.h))
        assert rc == 0, self.lib.pobrax_last_error()
        self.nb = self.lib.emu_num_bodies(self.h)

    def step(self, qp, act):
        a = [np.ascontiguousarray(x, np.float32).copy() for x in (qp.pos, qp.rot, qp.vel, qp.ang)]
        act = np.ascontiguousarray(act, np.float32)
        n = a[0].shape[0]
        assert a[0].shape == (n, self.nb, 3) and act.shape == (n, 8)
        cv, ca = np.zeros_like(a[0]), np.zeros_like(a[0])
        ptr = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = self.lib.emu_step(self.h, C.c_long(n), *[ptr(x) for x in a], ptr(act), ptr(cv), ptr(ca))
        assert rc == 0, rc
        return a, cv, ca

    def __del__(self):
        if getattr(self, 'h', None):
            self.lib.emu_destroy(self.h)


def _check(kind, oenv, emu, s, T, n, seed=1, sys_factory=None):
    """Teacher-forced: every env either meets the tight gates against the oracle step or (rounding-ambiguous
    decisions, tests/_parity.py TwoBranch) against the oracle step with some of those decisions taken the other
    way; the unexplained rest (<= 1 % of all env-steps) is held to the loose bound."""
    from oracle import threefry as tf
    oenv.sys.track_margin = True
    two = P.TwoBranch(sys_factory or (lambda: oenvs.ENVS[kind]().sys), threads=2)
    rng = tf.prng_key(seed)
    wall_hits, unexplained, marginal = 0, 0, 0
    for t in range(T):
        rng, act = P.actions_for(rng, n)
        oenv.sys.margin = None
        qp1, info = oenv.sys.step(s.qp, act)
        (pos, rot, vel, ang), cv, ca = emu.step(s.qp, act)
        got = P.Got(pos[:, :9], rot[:, :9], vel[:, :9], ang[:, :9], cv[:, :9], ca[:, :9])
        clear = oenv.sys.margin > P.BRANCH_MARGIN
        ok = P.physics_match(got, qp1, info, contact_clipped=False)
        assert ok[clear].all(), (kind, t, 'unambiguous envs off the tight gates', np.nonzero(clear & ~ok)[0][:8])
        rest = np.nonzero(~clear & ~ok)[0]
        expl, _ = two.explain(s.qp, act, rest, got, contact_clipped=False)
        bad = rest[~expl]
        marginal += int((~clear).sum()); unexplained += len(bad)
        for g, w, tol in ((pos, qp1.pos, P.LOOSE_POS), (rot, qp1.rot, P.LOOSE_POS), (vel, qp1.vel, P.LOOSE_VEL),
                          (ang, qp1.ang, P.LOOSE_VEL)):
            assert (np.abs(g[bad, :9] - w[bad, :9]) <= tol).all(), (kind, t, 'unexplained env off the loose bound')
        # frozen bodies are not part of the device state: untouched
        assert np.array_equal(pos[:, 9:], s.qp.pos[:, 9:])
        wall_hits += int((np.abs(info.contact_vel[:, [1, 3, 5, 7]]).sum(-1) > 0).sum())   # Aux bodies only touch walls
        s = s.replace(qp=qp1)
    assert unexplained <= P.MAX_UNEXPLAINED * n * T, (unexplained, marginal, n * T)
    return wall_hits


@pytest.mark.parametrize('kind', list(KINDS))
def test_device_substep_text_matches_oracle(kind):
    n, T = 48, 12
    oenv = oenvs.ENVS[kind]()
    s = oenv.reset(P.keys_for(n, seed=0))
    _check(kind, oenv, Emu(kind), s, T, n)


def test_device_wall_paths_match_oracle_at_the_corner():
    """HeavenHell ants spawned around the staircase corner of the T junction (as in the GPU corner test): cells
    with two candidate walls, Aux wall contacts, closest points inside the segment -- the out-of-line groups."""
    n, T = 64, 30
    box = ((1.3, 4.9), (2.0, 5.9))
    oenv = oenvs.ENVS['ant_heavenhell']()
    oenv._init_lo, oenv._init_hi = np.array(box[0], np.float32), np.array(box[1], np.float32)
    s = oenv.reset(P.keys_for(n, seed=0))
    hits = _check('ant_heavenhell', oenv, Emu('ant_heavenhell'), s, T, n)
    assert hits > 10, hits


def test_action_repeat_runs_more_substeps():
    """ActionRepeatWrapper (wrappers.py:16-24): dt and substeps scale, h stays."""
    n = 16
    oenv = oenvs.ENVS['ant'](action_repeat=2)
    s = oenv.reset(P.keys_for(n, seed=2))
    _check('ant', oenv, Emu('ant', action_repeat=2), s, 4, n, sys_factory=lambda: oenvs.ENVS['ant'](action_repeat=2).sys)


@pytest.mark.parametrize('kind', ['ant_heavenhell', 'ant_tag', 'ant_gather'])
def test_capsule_end_table_never_misses_a_wall_in_reach(kind):
    """The exact cull of the lower leg's wall contacts: whenever SOME point of the capsule's segment is within the
    capsule radius of a wall box, that wall must be listed at the tip's or at the knee's cell of the capsule-end
    table (faces: the distance along a segment is linear, smallest at an end; vertices: the nearer end is within half
    a segment). 200 000 random capsules thrown at the walls, incl. poses through and on top of them."""
    emu = Emu(kind)
    lib = emu.lib
    lo_hi = np.zeros((8, 6), np.float32)
    nw = lib.emu_walls(emu.h, lo_hi.ctypes.data_as(C.c_void_p))
    boxes = lo_hi[:nw]
    oenv = oenvs.ENVS[kind]()
    S = oenv.sys
    assert np.allclose(np.sort(boxes, axis=0), np.sort(S.boxes + np.array([0, 0, 0.5, 0, 0, 0.5], np.float32), axis=0), atol=1e-6)
    rng = np.random.default_rng(7)
    n, r = 200_000, 0.08
    k = list(S.cap_body).index(2)
    half = float(np.linalg.norm(S.cap_a[k] - S.cap_b[k])) / 2        # lower-leg half segment
    # centres near a random wall's boundary (within 0.6 m), random 3-D directions
    w = rng.integers(0, nw, n)
    u = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    c = boxes[w, :3] + u * (boxes[w, 3:] - boxes[w, :3])
    c[:, :2] += rng.normal(0, 0.35, (n, 2)).astype(np.float32)
    c[:, 2] = rng.uniform(0.0, 1.3, n)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d *= half / np.linalg.norm(d, axis=1, keepdims=True)
    a, b = (c + d).astype(np.float32), (c - d).astype(np.float32)
    sp, bp = S._closest_segment_box(a[:, None, :].repeat(nw, 1), b[:, None, :].repeat(nw, 1), boxes[None, :, :3], boxes[None, :, 3:])
    dist = np.sqrt(((sp - bp) ** 2).sum(-1))                          # [n, nw]
    touching = dist < r
    masks = np.zeros((2, n), np.uint32)
    for i, pts in enumerate((a, b)):
        xy = np.ascontiguousarray(pts[:, :2], np.float32)
        assert lib.emu_tip_masks(emu.h, C.c_long(n), xy.ctypes.data_as(C.c_void_p), masks[i].ctypes.data_as(C.c_void_p)) == 0
    listed = ((masks[0] | masks[1])[:, None] >> np.arange(nw)[None, :]) & 1
    assert touching.sum() > 5000                                       # the sample does exercise contacts
    assert not (touching & (listed == 0)).any(), np.argwhere(touching & (listed == 0))[:5]
    # and the table is worth having: most capsules that touch nothing are not flagged at all
    assert ((masks[0] | masks[1]) == 0)[~touching.any(1)].mean() > 0.3
