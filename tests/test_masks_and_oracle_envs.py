"""Host logic: the 'ant' observability masks, and env-level properties of the CPU oracle (reset sampling
rules and wrapper semantics restated from the reference)."""
import numpy as np
import torch

from oracle import envs as oenvs
from oracle import threefry as tf
from po_brax_b200 import standard_observability_masks as M


def test_ant_masks_are_the_reference_ranges():
    # /root/reference/po_brax/standard_observability_masks.py:7,26,62
    assert M.POSITION['ant'].tolist() == list(range(0, 13))
    assert M.VELOCITY['ant'].tolist() == list(range(13, 27))
    assert M.CFRC['ant'].tolist() == list(range(27, 87))
    assert sorted(M.POSITION['ant'].tolist() + M.VELOCITY['ant'].tolist() + M.CFRC['ant'].tolist()) == list(range(87))
    obs = np.arange(2 * 87, dtype=np.float32).reshape(2, 87)
    assert M.apply_mask(obs, M.VELOCITY['ant']).shape == (2, 14)
    assert torch.equal(M.apply_mask(torch.as_tensor(obs), M.POSITION['ant']), torch.as_tensor(obs[:, :13]))


def _keys(n, seed=0):
    return tf.split(tf.prng_key(seed), n + 1)[1:]


def test_oracle_observation_sizes():
    for name, d in (('ant', 87), ('ant_heavenhell', 114), ('ant_tag', 103), ('ant_gather', 211)):
        assert oenvs.ENVS[name]().reset(_keys(2)).obs.shape == (2, d)


def test_oracle_reset_rules():
    n = 64
    hh = oenvs.AntHeavenHellEnv()
    s = hh.reset(_keys(n))
    t, h = s.qp.pos[:, hh.target_idx], s.qp.pos[:, hh.hell_idx]
    assert set(np.unique(t[:, 0])) == {-5.25, 5.25} and (t[:, 0] == -h[:, 0]).all()      # one each side
    xy = s.qp.pos[:, 0, :2]
    assert (np.abs(xy[:, 0]) <= 0.5).all() and ((xy[:, 1] >= 0.5) & (xy[:, 1] <= 1.5)).all()
    assert np.array_equal(s.qp.pos[:, 9, :2], xy)                                        # Ground shifted too (:70)
    tag = oenvs.AntTagEnv()
    s = tag.reset(_keys(n))
    d = np.linalg.norm(s.qp.pos[:, tag.target_idx, :2] - s.qp.pos[:, 0, :2], axis=1)
    assert (d > 5.0).all() and (np.abs(s.qp.pos[:, tag.target_idx, :2]) <= 4.5).all()
    assert tag.last_reject_iters.max() > 0
    g = oenvs.AntGatherEnv()
    s = g.reset(_keys(n))
    assert len(g.grid) == 156 and (g.waiting_area == [18, 18, 12]).all()
    idx = g.last_object_idx
    assert all(len(set(r)) == 16 for r in idx.tolist())                                   # without replacement
    assert (s.qp.pos[:, g.obj[:8], 2] == 1.0).all() and (s.qp.pos[:, g.obj[8:], 2] == 0.0).all()


def test_oracle_episode_and_autoreset():
    n, L = 8, 3
    env = oenvs.create('ant', episode_length=L, auto_reset=True)
    s = env.reset(_keys(n))
    first = s.qp.pos.copy()
    a = np.zeros((n, 8), np.float32)
    for t in range(1, 2 * L + 1):
        s = env.step(s, a)
        want_steps = ((t - 1) % L) + 1
        assert (s.info['steps'] == want_steps).all()
        if want_steps == L:
            assert (s.done == 1).all() and (s.info['truncation'] == 1).all()
            assert np.array_equal(s.qp.pos, first)       # cached first state restored
        else:
            assert (s.done == 0).all() and (s.info['truncation'] == 0).all()


def test_host_threefry_matches_oracle():
    from po_brax_b200 import random as prandom
    for seed in (0, 1, 12345, 2 ** 33 + 7):
        k = prandom.prng_key(seed)
        assert list(k) == tf.prng_key(seed).tolist()
        full = tf.split(tf.prng_key(seed), 9)
        for j in (0, 1, 4, 8):
            assert list(prandom.split_at(k, 9, j)) == full[j].tolist()
    assert prandom.threefry2x32((0x13198a2e, 0x03707344), 0x243f6a88, 0x85a308d3) == (0xc4923a9c, 0x483df7a0)


def test_bench_flop_figure_is_not_inflated():
    """bench.py's roofline uses SURVEY 8(d)'s specialised count (4.30e4 FLOPs per env-step). The oracle's own
    System.step, run on operation-counting arrays (oracle/count_ops.py), is the as-written upper bound: the bench figure
    must not exceed it, and the structural counts must hold (8 joints x 10 substeps = 80 atan2)."""
    import bench
    from oracle import count_ops
    c = count_ops.count_step('ant', n=2)
    total = count_ops.flops(c)
    assert c['atan2'] == 80 and not any(k.startswith('other:') for k in c)
    assert bench.FLOPS_PER_ENV_STEP['ant'] <= total <= 1.5 * bench.FLOPS_PER_ENV_STEP['ant']
    hh = count_ops.flops(count_ops.count_step('ant_heavenhell', walls=False, n=2))
    assert bench.FLOPS_PER_ENV_STEP['ant_heavenhell'] <= hh
