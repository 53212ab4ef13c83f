"""Pins oracle/brax_v1.py against the only numeric fixture the reference ships: the 21-frame Ant-Tag
rollout embedded in /root/reference/notebooks/ant_tag.ipynb:449 (extracted by tests/golden/make_fixture.py).

The notebook ran un-jitted reset + jitted steps with a NumPy key, i.e. brax-jumpy's NumPy-PCG64 branch
for sampling; the sampling below restates that branch, the physics is the oracle's."""
import json
import os

import numpy as np
import pytest

from oracle import brax_v1 as bx

HERE = os.path.dirname(os.path.abspath(__file__))


def _split_np(rng, n):
    return np.random.default_rng(rng).integers(0, 2 ** 32, dtype=np.uint32, size=(n, 2))


def _uniform_np(rng, n, lo, hi):
    return np.random.default_rng(rng).uniform(lo, hi, n)


def _replay(dtype, c_step=False):
    cfg = json.load(open(os.path.join(HERE, 'golden', 'ant_tag_config.json')))
    sys_ = bx.System(cfg, dtype=dtype, walls=False)  # that revision's Arena: capsule walls at +-7, never reached
    if c_step:  # the scalar C restatement (oracle/brax_step.c) replays the same fixture
        from oracle import cstep
        cstep.attach(sys_, threads=1)
    rng = np.random.default_rng(0).integers(0, 2 ** 32, dtype=np.uint32, size=2)
    assert rng.tolist() == [3653403231, 2735729615]
    ks = _split_np(rng, 5)
    qpos = sys_.default_angle() + _uniform_np(ks[1], 8, -.1, .1).astype(dtype)
    qvel = _uniform_np(ks[2], 8, -.1, .1).astype(dtype)
    qp = sys_.default_qp(qpos[None].astype(dtype), qvel[None].astype(dtype))
    xy = _uniform_np(ks[1], 2, -4.5, 4.5).astype(dtype)
    qp.pos[:, :10, :2] += xy
    frames = [(qp.pos[0].copy(), qp.rot[0].copy())]
    for _ in range(20):
        rng, rng1 = _split_np(rng, 2)
        act = _uniform_np(rng1, 8, -1, 1).astype(np.float32).astype(dtype)
        qp, _info = sys_.step(qp, act[None])
        frames.append((qp.pos[0].copy(), qp.rot[0].copy()))
    return frames


@pytest.mark.parametrize('c_step', [False, True], ids=['numpy', 'c'])
@pytest.mark.parametrize('dtype,tol0,tol1,tol20', [(np.float64, 1e-7, 2e-6, 3e-5), (np.float32, 1e-6, 2e-6, 3e-5)])
def test_golden_rollout(dtype, tol0, tol1, tol20, c_step):
    gold = np.load(os.path.join(HERE, 'golden', 'ant_tag_rollout.npz'))
    frames = _replay(dtype, c_step)
    errs = []
    for t, (pos, rot) in enumerate(frames):
        e = max(np.abs(pos[:9] - gold['pos'][t, :9]).max(), np.abs(rot[:9] - gold['rot'][t, :9]).max())
        errs.append(e)
    assert errs[0] <= tol0, errs[0]
    assert errs[1] <= tol1, errs[1]
    assert max(errs) <= tol20, errs


def test_default_qp_torso_height():
    gold = np.load(os.path.join(HERE, 'golden', 'ant_tag_rollout.npz'))
    frames = _replay(np.float32)
    assert abs(frames[0][0][0, 2] - 0.5365347) < 1e-6
    assert abs(gold['pos'][0, 0, 2] - 0.5365347) < 1e-6
