"""Pins oracle/brax_step.c (the scalar C / OpenMP restatement that bench.py times as the CPU baseline) to the
NumPy oracle oracle/brax_v1.py, which tests/test_oracle_golden.py pins to the reference's notebook fixture (the C
text replays that fixture too, see test_golden_rollout[...-c]).

  * float64 build vs NumPy float64, teacher-forced, all four systems (ground + wall contacts, actuators):
    |d| <= 1e-10 on qp and contact impulses -- same expressions in the same order, only libm's atan2 may differ.
  * float32 build vs NumPy float32: the tolerances of tests/_parity.py (pos/rot 1e-6 + 2e-6|x|, vel/ang 3e-4) on
    the envs whose contact / actuator decisions are not rounding-ambiguous in that step (oracle's own margin report).
  * thread count does not change results (envs are independent).
No GPU, no po_brax_b200 import: this is oracle-vs-oracle."""
import numpy as np
import pytest

from oracle import cstep, envs as oenvs, threefry as tf
from tests._parity import BRANCH_MARGIN, POS_TOL, VEL_ATOL

NAMES = ['ant', 'ant_heavenhell', 'ant_tag', 'ant_gather']


def _pair(name, dtype, threads=2):
    a, b = oenvs.ENVS[name](dtype=dtype), oenvs.ENVS[name](dtype=dtype)
    cstep.attach(b.sys, threads=threads)
    return a, b


def _fields(qp, info):
    return dict(pos=qp.pos, rot=qp.rot, vel=qp.vel, ang=qp.ang, cvel=info.contact_vel, cang=info.contact_ang)


@pytest.mark.parametrize('name', NAMES)
def test_c_step_matches_numpy_f64(name):
    n, T = 96, 25
    ref, cenv = _pair(name, np.float64)
    keys = tf.split(tf.prng_key(11), n + 1)[1:]
    s, sc = ref.reset(keys), cenv.reset(keys)
    assert np.array_equal(s.obs, sc.obs)          # reset goes through System.info in C
    rng = np.random.default_rng(5)
    qp = s.qp
    for t in range(T):
        act = rng.uniform(-1, 1, (n, 8))
        q1, i1 = ref.sys.step(qp, act)
        q2, i2 = cenv.sys.step(qp, act)
        for k, (x, y) in ((k, (v, _fields(q2, i2)[k])) for k, v in _fields(q1, i1).items()):
            assert np.abs(x - y).max() <= 1e-10, (name, t, k, np.abs(x - y).max())
        qp = q1


@pytest.mark.parametrize('name', NAMES)
def test_c_step_matches_numpy_f32(name):
    n, T = 128, 25
    ref, cenv = _pair(name, np.float32)
    ref.sys.track_margin = True
    keys = tf.split(tf.prng_key(12), n + 1)[1:]
    qp = ref.reset(keys).qp
    rng = np.random.default_rng(6)
    compared = 0
    for t in range(T):
        act = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        ref.sys.margin = None
        q1, i1 = ref.sys.step(qp, act)
        q2, i2 = cenv.sys.step(qp, act)
        ok = ref.sys.margin > BRANCH_MARGIN
        compared += int(ok.sum())
        f1, f2 = _fields(q1, i1), _fields(q2, i2)
        for k in ('pos', 'rot'):
            d = np.abs(f1[k][ok] - f2[k][ok])
            assert (d <= POS_TOL[0] + POS_TOL[1] * np.abs(f1[k][ok])).all(), (name, t, k, d.max())
        for k in ('vel', 'ang', 'cvel', 'cang'):
            d = np.abs(f1[k][ok] - f2[k][ok])
            assert d.max() <= VEL_ATOL, (name, t, k, d.max())
        qp = q1
    assert compared >= 0.75 * n * T, compared


def test_threads_do_not_change_results():
    n = 64
    e1, e4 = oenvs.ENVS['ant_heavenhell'](), oenvs.ENVS['ant_heavenhell']()
    cstep.attach(e1.sys, threads=1)
    cstep.attach(e4.sys, threads=4)
    keys = tf.split(tf.prng_key(1), n + 1)[1:]
    s1, s4 = e1.reset(keys), e4.reset(keys)
    act = tf.uniform(tf.prng_key(2), n * 8, -1.0, 1.0).reshape(n, 8)
    for _ in range(5):
        s1, s4 = e1.step(s1, act), e4.step(s4, act)
    for x, y in ((s1.qp.pos, s4.qp.pos), (s1.qp.vel, s4.qp.vel), (s1.obs, s4.obs), (s1.reward, s4.reward)):
        assert np.array_equal(x, y)


def test_wrapped_env_runs_on_c_step():
    """The configuration bench.py times: create() = Episode + cached AutoReset on top of the C step."""
    n = 32
    env, cenv = oenvs.create('ant_heavenhell'), oenvs.create('ant_heavenhell')
    cstep.attach(cenv.env.sys, threads=2)
    keys = tf.split(tf.prng_key(3), n + 1)[1:]
    s, sc = env.reset(keys), cenv.reset(keys)
    act = tf.uniform(tf.prng_key(4), n * 8, -1.0, 1.0).reshape(n, 8)
    for _ in range(3):
        s, sc = env.step(s, act), cenv.step(sc, act)
    assert np.array_equal(s.done, sc.done) and np.array_equal(s.info['steps'], sc.info['steps'])
    assert np.median(np.abs(s.qp.pos - sc.qp.pos)) <= 1e-6
