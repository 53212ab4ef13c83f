"""CUDA path vs the CPU oracle through the public env API (which calls the C-ABI via ctypes).
Run on the B200 box: python -m pytest tests -m gpu."""
import os

import numpy as np
import pytest
import torch

from oracle import envs as oenvs
from oracle import threefry as tf
from tests import _parity as P

pytestmark = pytest.mark.gpu

KINDS = ['ant', 'ant_heavenhell', 'ant_tag', 'ant_gather']


def _make(kind, n, **kw):
    from po_brax_b200 import envs
    return envs.create(kind, batch_size=n, **kw)


def _oracle_aux(kind, oenv, s):
    """The per-env frozen-body data the CUDA path keeps in `aux`, from the oracle's full qp."""
    if kind == 'ant_heavenhell':
        return s.qp.pos[:, oenv.target_idx, 0]
    if kind == 'ant_tag':
        return s.qp.pos[:, oenv.target_idx]
    if kind == 'ant_gather':
        return s.qp.pos[:, oenv.obj]
    return None


@pytest.mark.parametrize('kind', KINDS)
def test_reset_parity(kind):
    n = 256
    keys = P.keys_for(n, seed=0)
    oenv = oenvs.ENVS[kind]()
    want = oenv.reset(keys)
    env = _make(kind, n)
    got = env.reset(keys)
    torch.cuda.synchronize()
    nb = oenv.sys.num_bodies
    assert env.observation_size == want.obs.shape[1] and env.num_bodies == nb
    # integers / RNG: bit-exact
    if kind != 'ant':
        assert (P.rng_bits(got.info['rng']) == want.info['rng']).all()
    gq = got.qp
    if kind == 'ant_heavenhell':
        assert (P.t2n(gq.pos)[:, oenv.target_idx] == want.qp.pos[:, oenv.target_idx]).all()  # heaven side
        assert (P.t2n(gq.pos)[:, oenv.hell_idx] == want.qp.pos[:, oenv.hell_idx]).all()
        assert (P.t2n(gq.pos)[:, oenv.priest_idx] == want.qp.pos[:, oenv.priest_idx]).all()
    if kind == 'ant_tag':
        assert (P.t2n(gq.pos)[:, oenv.target_idx] == want.qp.pos[:, oenv.target_idx]).all()  # rejection loop
    if kind == 'ant_gather':
        assert (P.t2n(gq.pos)[:, oenv.obj] == want.qp.pos[:, oenv.obj]).all()                # 16-of-156 choice
    # every frozen body (Ground, Arena, Priest/Target/Hell, apples/bombs), exactly
    assert np.array_equal(P.t2n(gq.pos)[:, 9:], want.qp.pos[:, 9:])
    assert np.array_equal(P.t2n(gq.rot)[:, 9:], want.qp.rot[:, 9:])
    P.assert_qp_close(gq, want.qp, f'{kind} reset')
    # obs: the contact columns of a body resting exactly on z = 0 (|pen| ~ 1 ulp after the z-lift) flip with
    # rounding in the reference itself; they are compared only where the oracle's penetration is unambiguous
    gobs = P.t2n(got.obs)
    mask = np.ones_like(want.obs, bool)
    Pc = 1 if kind == 'ant' else 3
    cv0 = Pc + 26
    ambiguous = _ambiguous_contacts(oenv, want.qp)
    for b in range(9):
        cols = [cv0 + 3 * b + c for c in range(3)] + [cv0 + 3 * nb + 3 * b + c for c in range(3)]
        mask[np.ix_(ambiguous[:, b], cols)] = False
    P.assert_obs_close(gobs, want.obs, kind, nb, f'{kind} reset', mask=mask)
    assert mask.mean() > 0.85  # at most one resting foot (6 contact columns) per env is excluded
    for name in ('reward', 'done'):
        assert (P.t2n(getattr(got, name)) == 0).all()
    assert (P.t2n(got.info['steps']) == 0).all()


def _ambiguous_contacts(oenv, qp, eps=1e-5, walls_only=False):
    """[N, 9] bool: bodies with a ground-contact candidate whose |penetration| < eps (sign decided by rounding),
    or (walls) a capsule within eps of touching a wall. walls_only + eps=inf: bodies whose capsule is within its
    radius of a wall box (touching)."""
    from oracle import brax_v1 as bx
    s = oenv.sys
    n = qp.pos.shape[0]
    amb = np.zeros((n, 9), bool)
    b = s.cp_body
    end_w = qp.pos[:, b] + bx.rotate(np.broadcast_to(s.cp_end, qp.pos[:, b].shape), qp.rot[:, b])
    pen = -(end_w[..., 2] - s.cp_rad)
    if not walls_only:
        for k, body in enumerate(b):
            amb[:, body] |= np.abs(pen[:, k]) < eps
    if len(s.boxes):
        nbx = len(s.boxes)
        bb = np.repeat(s.cap_body, nbx)
        ca, cb = np.repeat(s.cap_a, nbx, axis=0), np.repeat(s.cap_b, nbx, axis=0)
        rad = np.repeat(s.cap_rad, nbx)
        box = np.tile(s.boxes, (len(s.cap_body), 1))
        pos, rot = qp.pos[:, bb], qp.rot[:, bb]
        apos = qp.pos[:, s.arena][:, None, :]
        a_w = pos + bx.rotate(np.broadcast_to(ca, pos.shape), rot)
        b_w = pos + bx.rotate(np.broadcast_to(cb, pos.shape), rot)
        sp, bp = s._closest_segment_box(a_w, b_w, apos + box[:, :3], apos + box[:, 3:])
        d = np.sqrt(((sp - bp) ** 2).sum(-1))
        near = (np.abs(rad - d) < eps) if np.isfinite(eps) else ((d < rad) & (d > 0))
        for k, body in enumerate(bb):
            amb[:, body] |= near[:, k]
    return amb


@pytest.mark.parametrize('kind', KINDS)
def test_step_teacher_forced(kind):
    """One env step from identical states, T times along an oracle rollout (BASELINE config 1 key scheme)."""
    _teacher_forced(kind, 128, 25)


def test_step_teacher_forced_long_heavenhell():
    """BASELINE config 1 (Ant-HeavenHell, 128 envs, 1000-step random-action rollout), teacher-forced over the whole
    episode length (ants against walls, fallen ants, finished envs). POBRAX_LONG_T shortens it for quick runs."""
    _teacher_forced('ant_heavenhell', 128, int(os.environ.get('POBRAX_LONG_T', '1000')), c_step=True)


def test_step_teacher_forced_corner_walls():
    """HeavenHell ants spawned around the staircase corner of the T junction (wall boxes x in [2, 3] up to y = 5.5 and
    y in [5, 6] from x = 2.5): two candidate walls per cell, Aux / torso wall contacts, closest points inside the
    segment (bisection) -- the out-of-line wall groups and the multi-candidate cull against the oracle."""
    st = _teacher_forced('ant_heavenhell', 128, 40, init_box=((1.3, 4.9), (2.0, 5.9)))
    assert st['aux_wall_contacts'] > 20 and st['torso_contacts'] > 0, st   # the scenario does reach those colliders


@pytest.mark.parametrize('kind,n,T', [('ant', 4096, 6), ('ant_gather', 16384, 4), ('ant_tag', 65536, 3),
                                      ('ant_heavenhell', 131072, 3)])
def test_step_teacher_forced_at_baseline_sizes(kind, n, T):
    """BASELINE configs 2-4 at their full batch sizes and config 5 at its 8-GPU shard size (1 Mi / 8), teacher-forced
    against the scalar C twin of the oracle (oracle/brax_step.c on all host cores; tests/test_oracle_c.py pins it to
    the NumPy text and to the golden rollout) under the NumPy task logic: every env of the batch is compared, and at
    most 1 % of them may be 'unexplained' (tests/_parity.py TwoBranch)."""
    st = _teacher_forced(kind, n, T, c_step=True)
    print(f'\n[parity] {kind} n={n} T={T}: {st}')


@pytest.mark.parametrize('kind,n', [('ant_heavenhell', 32768), ('ant_tag', 32768)])
def test_step_teacher_forced_ants_piled_up_along_the_walls(kind, n):
    """The wall paths at scale: the oracle first walks the batch 250 steps under a PERIODIC action sequence (a gait:
    the ants travel metres and pile up along the walls -- 15-20 % of the envs touch one in any step, against 0-12 %
    from a reset), then every env is teacher-forced as usual. Exercises the capsule-end tables, `tip_wall`, the
    out-of-line group and fallen ants (torso on the ground) on tens of thousands of envs."""
    st = _teacher_forced(kind, n, 3, c_step=True, preroll=(250, 4))
    print(f'\n[parity] {kind} n={n} after a 250-step gait: {st}')
    assert st['wall_contact_envs'] > 0.05 * n, st


def _teacher_forced(kind, n, T, init_box=None, c_step=False, preroll=None):
    """One env step from identical states, T times along an oracle rollout. Every env must meet the TIGHT gates of
    tests/_parity.py against the oracle's step -- or, where the oracle reports a rounding-ambiguous contact / actuator
    decision in that step, against the oracle's step with some of those decisions taken the other way (two-branch
    check). Envs that match neither are 'unexplained': loose bound, and at most 1 % of the batch in any step."""
    keys = P.keys_for(n, seed=0)
    oenv = oenvs.ENVS[kind]()
    if c_step:
        from oracle import cstep
        cstep.attach(oenv.sys, threads=os.cpu_count() or 1)
    two = P.TwoBranch(lambda: oenvs.ENVS[kind]().sys)
    nb = oenv.sys.num_bodies
    kw = {}
    if init_box is not None:   # ant_heavenhell.py:73 self._init_ant_pos
        oenv._init_lo, oenv._init_hi = np.array(init_box[0], np.float32), np.array(init_box[1], np.float32)
        kw['init_ant_pos'] = init_box
    s = oenv.reset(keys)
    if preroll is not None:   # physics-only walk under a periodic action sequence (the task state stays as reset)
        steps, period = preroll
        gait = np.random.default_rng(3).uniform(-1, 1, (period, n, 8)).astype(np.float32)
        qp = s.qp
        for t in range(steps):
            qp, _ = oenv.sys.step(qp, gait[t % period])
        s = s.replace(qp=qp)
    env = _make(kind, n, auto_reset=False, episode_length=1000, **kw)
    rng = tf.prng_key(1)
    oenv.sys.track_margin = True
    stats = {'aux_wall_contacts': 0, 'torso_contacts': 0, 'marginal': 0.0, 'other_branch': 0, 'unexplained': 0,
             'worst_unexplained_share': 0.0, 'wall_contact_envs': 0}
    Pc = 1 if kind == 'ant' else 3
    for t in range(T):
        rng, a = P.actions_for(rng, n)
        cs = env.state_from_qp(P.qp_to_torch(s.qp), rng=s.info.get('rng'))
        oenv.sys.margin = None
        nxt = oenv.step(oenvs.State(s.qp.copy(), s.obs, s.reward, s.done, dict(s.metrics), dict(s.info)), a)
        got = env.step(cs, torch.as_tensor(a, device='cuda'))
        torch.cuda.synchronize()
        # Envs in which some contact sat within rounding noise of a discontinuous branch of the reference
        # algorithm (touching / approaching / J > 0 / |v_d| > 0.01 / actuator cut-off; oracle/brax_v1.py:_note_margin)
        # may take the other branch under any other float32 evaluation order (XLA's included).
        clear = oenv.sys.margin > P.BRANCH_MARGIN
        g = P.Got.from_state(got, kind, nb)
        base_ok = P.physics_match(g, nxt.qp, bx_info_from_obs(nxt.obs, Pc, nb))
        assert base_ok[clear].all(), (f'{kind} t={t}: {int((clear & ~base_ok).sum())} unambiguous envs off the tight '
                                      f'gates, first {np.nonzero(clear & ~base_ok)[0][:6].tolist()}')
        rest = np.nonzero(~clear & ~base_ok)[0]
        explained, _ = two.explain(s.qp, a, rest, g)
        bad = rest[~explained]
        stats['marginal'] += float((~clear).mean()) / T
        stats['other_branch'] += int(explained.sum())
        stats['unexplained'] += len(bad)
        stats['worst_unexplained_share'] = max(stats['worst_unexplained_share'], len(bad) / n)
        assert len(bad) <= max(1, P.MAX_UNEXPLAINED * n), f'{kind} t={t}: {len(bad)} of {n} envs match neither branch'
        cvel = nxt.obs[:, Pc + 26:Pc + 26 + 3 * nb].reshape(n, nb, 3)   # clip(contact.vel): Aux bodies only touch walls
        stats['aux_wall_contacts'] += int((np.abs(cvel[:, [1, 3, 5, 7]]).sum(-1) > 0).sum())
        stats['torso_contacts'] += int((np.abs(cvel[:, 0]).sum(-1) > 0).sum())
        if t == 0 and len(oenv.sys.boxes):   # envs with some capsule within its radius of a wall box right now
            stats['wall_contact_envs'] = int(_ambiguous_contacts(oenv, s.qp, eps=np.inf, walls_only=True).any(1).sum())
        same = np.ones(n, bool)      # envs on the oracle's own branch: everything below is compared against `nxt`
        same[rest] = False
        P.assert_qp_close(got.qp, nxt.qp, f'{kind} t={t}', rows=same)
        P.assert_qp_close(got.qp, nxt.qp, f'{kind} t={t} (unexplained envs)', rows=bad, loose=True)
        assert np.array_equal(P.t2n(got.done), np.asarray(nxt.done, np.float32)), f'{kind} t={t} done'
        if kind == 'ant':
            # forward = dx / dt amplifies the pos tolerance x20
            assert (np.abs(P.t2n(got.reward) - nxt.reward)[same] <= 2e-4).all(), f'{kind} t={t} reward'
        else:
            assert np.array_equal(P.t2n(got.reward), nxt.reward), f'{kind} t={t} reward'
        if kind == 'ant_tag':
            assert (P.rng_bits(got.info['rng']) == nxt.info['rng']).all()
            assert np.array_equal(P.t2n(got.metrics['hits']), nxt.metrics['hits'])
            # the opponent moves along (ant - target)/|ant - target| of the post-physics ant: float tolerance
            # (tight where the ant's own step took the oracle's branch; an env on the other contact branch moves
            # its opponent along a slightly different direction)
            dtgt = np.abs(P.t2n(got.qp.pos)[:, oenv.target_idx] - nxt.qp.pos[:, oenv.target_idx]).max(axis=-1)
            assert dtgt[same].max() <= 1e-5 and dtgt.max() <= P.LOOSE_POS, (dtgt[same].max(), dtgt.max())
        if kind == 'ant_gather':
            assert np.array_equal(P.t2n(got.metrics['apples']), nxt.metrics['apples'].astype(np.float32))
            assert np.array_equal(P.t2n(got.metrics['bombs']), nxt.metrics['bombs'].astype(np.float32))
            assert np.array_equal(P.t2n(got.qp.pos)[:, oenv.obj], nxt.qp.pos[:, oenv.obj])
        mask = np.ones_like(nxt.obs, bool)
        mask[~same] = False       # other-branch envs: qp and contact columns were checked against their own branch
        if kind == 'ant_gather':  # a sensor bin index is int(trunc(angle / res)): tolerate angles on a bin edge
            mask[:, -2 * oenv.n_bins:] &= _gather_reading_mask(oenv, nxt, got)
        P.assert_obs_close(P.t2n(got.obs), nxt.obs, kind, nb, f'{kind} t={t}', mask=mask)
        s = nxt
    return stats


class bx_info_from_obs:
    """Info-like view of the oracle observation's (already clipped) contact columns, for P.physics_match."""

    def __init__(self, obs, Pc, nb):
        n = obs.shape[0]
        c0 = Pc + 26
        self.contact_vel = obs[:, c0:c0 + 3 * nb].reshape(n, nb, 3)
        self.contact_ang = obs[:, c0 + 3 * nb:c0 + 6 * nb].reshape(n, nb, 3)


def _gather_reading_mask(oenv, nxt, got):
    g, w = P.t2n(got.obs)[:, -2 * oenv.n_bins:], nxt.obs[:, -2 * oenv.n_bins:]
    ok = np.abs(g - w) <= 1e-5
    # rows that differ are accepted only if they hold the same multiset of intensities shifted by one bin
    bad_rows = ~ok.all(axis=1)
    mask = np.ones_like(ok)
    for r in np.nonzero(bad_rows)[0]:
        if np.allclose(np.sort(g[r]), np.sort(w[r]), atol=1e-5):
            mask[r] = False
    assert bad_rows.mean() <= 0.01
    return mask


@pytest.mark.parametrize('kind', ['ant_heavenhell', 'ant_tag'])
def test_free_running_20_steps(kind):
    """Free-running rollout, 20 steps (chaos sets in later). Gate over EVERY env whose 20 steps stayed clear of a
    rounding-ambiguous decision (the oracle's own margins): max |d pos|, |d rot| <= 1e-4, and the SURVEY App. C figure
    (3e-5, measured there on one env) for >= 95 % of them. Basis: the device text under g++ (tests/host_emu) is
    2.1e-5 / 3.6e-5 (HeavenHell / Tag) at worst over ~220 clear envs after 20 steps, 90th percentile 0.8e-5 / 1.3e-5
    -- errors grow ~3x per 5 steps, so a hard max over hundreds of envs sits above a single env's figure. An env
    that did pass through an ambiguous decision may legitimately be one contact impulse apart: not gated here (the
    teacher-forced two-branch tests cover those steps)."""
    n, T = 256, 20
    keys = P.keys_for(n, seed=3)
    oenv = oenvs.ENVS[kind]()
    from oracle import cstep
    cstep.attach(oenv.sys, threads=os.cpu_count() or 1)
    oenv.sys.track_margin = True
    s = oenv.reset(keys)
    env = _make(kind, n, auto_reset=False)
    cs = env.reset(keys)
    rng = tf.prng_key(1)
    dirty = np.zeros(n, bool)
    for t in range(T):
        rng, a = P.actions_for(rng, n)
        oenv.sys.margin = None
        s = oenv.step(s, a)
        dirty |= oenv.sys.margin <= P.BRANCH_MARGIN
        cs = env.step(cs, torch.as_tensor(a, device='cuda'))
    q = cs.qp
    err = np.maximum(np.abs(P.t2n(q.pos)[:, :9] - s.qp.pos[:, :9]).reshape(n, -1).max(1),
                     np.abs(P.t2n(q.rot)[:, :9] - s.qp.rot[:, :9]).reshape(n, -1).max(1))
    clear = ~dirty
    assert clear.mean() > 0.25, clear.mean()
    print(f'\n[parity] free-running {kind}: {int(clear.sum())} clear envs, max {err[clear].max():.2e}, '
          f'q95 {np.quantile(err[clear], 0.95):.2e}, median {np.median(err[clear]):.2e}')
    assert err[clear].max() <= 1e-4, np.sort(err[clear])[-8:]
    assert np.quantile(err[clear], 0.95) <= 3e-5, np.sort(err[clear])[-8:]
    assert np.median(err[clear]) <= 1e-5


@pytest.mark.parametrize('kind,n,T', [('ant_heavenhell', 4096, 400), ('ant_tag', 4096, 300), ('ant_gather', 2048, 300),
                                      ('ant', 2048, 300)])
def test_free_running_rollout_statistics(kind, n, T):
    """Past ~100 steps a float32 rollout is chaotic, so env-by-env comparison is meaningless -- but the two
    implementations must still sample the same process. Free-running create(...) envs (Episode + cached AutoReset)
    on identical keys and actions for T steps, CUDA vs the C twin of the oracle under the NumPy task logic: the
    per-env totals of reward, finished episodes and (HeavenHell) each outcome, and the final torso height, must have
    equal means within 5 standard errors of a two-sample test (conservative: the samples are positively correlated)."""
    from oracle import cstep
    keys = P.keys_for(n, seed=5)
    oenv = oenvs.create(kind)
    cstep.attach(oenv.env.sys, threads=os.cpu_count() or 1)
    env = _make(kind, n)
    s, cs = oenv.reset(keys), env.reset(keys)
    names = ['reward', 'done'] + (['heaven', 'hell', 'dead'] if kind == 'ant_heavenhell' else [])

    def cats(r, d, xp):
        out = [r, d]
        if kind == 'ant_heavenhell':
            out += [(r == 1).astype(np.float32) if xp is np else (r == 1).float(),
                    (r == -1).astype(np.float32) if xp is np else (r == -1).float(),
                    (r == -2).astype(np.float32) if xp is np else (r == -2).float()]
        return out

    tot_o = [np.zeros(n, np.float64) for _ in names]
    tot_g = [torch.zeros(n, dtype=torch.float64, device='cuda') for _ in names]
    rng = tf.prng_key(2)
    for t in range(T):
        rng, a = P.actions_for(rng, n)
        s = oenv.step(s, a)
        cs = env.step(cs, torch.as_tensor(a, device='cuda'))
        for acc, v in zip(tot_o, cats(np.asarray(s.reward, np.float32), np.asarray(s.done, np.float32), np)):
            acc += v
        for acc, v in zip(tot_g, cats(cs.reward, cs.done.float(), torch)):
            acc += v
        if t == 19:   # still in the deterministic regime: the same envs have finished an episode
            assert (P.t2n(tot_g[1]) != tot_o[1]).sum() <= n // 1000
    rows = {k: (P.t2n(g), o) for k, g, o in zip(names, tot_g, tot_o)}
    rows['torso_z'] = (P.t2n(cs.qp.pos)[:, 0, 2].astype(np.float64), s.qp.pos[:, 0, 2].astype(np.float64))
    report = {}
    for k, (g, o) in rows.items():
        assert np.isfinite(g).all() and np.isfinite(o).all(), k
        se = np.sqrt((g.var() + o.var()) / n)
        report[k] = (float(g.mean()), float(o.mean()), float(se))
        assert abs(g.mean() - o.mean()) <= 5 * se + 1e-6, (kind, k, report[k])
    if kind == 'ant_heavenhell':
        assert report['done'][0] > 0.02, report   # the rollout is long enough to finish episodes (7.6 % of the envs)
    print(kind, report)


@pytest.mark.parametrize('kind', KINDS)
def test_episode_and_cached_autoreset(kind):
    """brax EpisodeWrapper + AutoResetWrapper semantics fused into the step kernel (create(auto_reset=True))."""
    n, L, T = 64, 4, 11
    keys = P.keys_for(n, seed=5)
    oenv = oenvs.create(kind, episode_length=L, auto_reset=True)
    s = oenv.reset(keys)
    env = _make(kind, n, episode_length=L, auto_reset=True)
    cs = env.reset(keys)
    rng = tf.prng_key(2)
    oenv.env.sys.track_margin = True
    dirty = np.zeros(n, bool)  # envs whose current episode went through a rounding-ambiguous branch (see _parity.py)
    for t in range(T):
        rng, a = P.actions_for(rng, n)
        oenv.env.sys.margin = None
        s = oenv.step(s, a)
        cs = env.step(cs, torch.as_tensor(a, device='cuda'))
        dirty |= oenv.env.sys.margin <= P.BRANCH_MARGIN
        assert np.array_equal(P.t2n(cs.info['steps']), s.info['steps']), f't={t}'
        ok = ~dirty
        assert np.array_equal(P.t2n(cs.done)[ok], np.asarray(s.done, np.float32)[ok]), f't={t}'
        assert np.array_equal(P.t2n(cs.info['truncation'])[ok], s.info['truncation'][ok]), f't={t}'
        # within an episode of 4 steps free-running drift stays far below the gates; after a reset both sides
        # return to the cached first state
        P.assert_qp_close(cs.qp, s.qp, f'{kind} autoreset t={t}', vel_atol=5e-3, pos_scale=10.0, rows=ok)
        dirty &= ~np.asarray(s.done, bool)  # episode over: both sides are back on the cached first state
    assert (~dirty).mean() > 0.5


@pytest.mark.parametrize('kind', ['ant_heavenhell', 'ant_gather', 'ant_tag'])
def test_gym_reset_where_done(kind):
    """wrappers.py:245-262: where done, qp/obs <- fresh reset(keys[i]); steps <- 0; everything else kept."""
    n = 64
    keys0, keys1 = P.keys_for(n, seed=7), P.keys_for(n, seed=8)
    env = _make(kind, n, auto_reset=False, episode_length=3)
    cs = env.reset(keys0)
    a = torch.zeros((n, 8), device='cuda')
    for _ in range(2):
        cs = env.step(cs, a)
    done = torch.zeros(n, device='cuda')
    done[::3] = 1.0
    cs.buf['done'].copy_(done)
    before = {k: v.clone() for k, v in cs.buf.items() if v is not None}
    fresh = env.reset(keys1)
    cs = env.reset_where_done(cs, keys1)
    torch.cuda.synchronize()
    d = done.bool()
    assert torch.equal(cs.buf['qp'][:, d], fresh.buf['qp'][:, d]) and torch.equal(cs.buf['qp'][:, ~d], before['qp'][:, ~d])
    assert torch.equal(cs.obs[d], fresh.obs[d]) and torch.equal(cs.obs[~d], before['obs'][~d])
    assert torch.equal(cs.buf['aux'][:, d], fresh.buf['aux'][:, d]) and torch.equal(cs.buf['aux'][:, ~d], before['aux'][:, ~d])
    assert (cs.info['steps'][d] == 0).all() and torch.equal(cs.info['steps'][~d], before['steps'][~d])
    for k in ('reward', 'done', 'truncation', 'rng', 'metrics'):
        assert torch.equal(cs.buf[k], before[k]), k


@pytest.mark.parametrize('kind', KINDS)
def test_pack_unpack_roundtrip_and_split_keys(kind):
    n = 96
    env = _make(kind, n)
    cs = env.reset(P.keys_for(n, seed=9))
    q = cs.qp
    qp2, aux2 = env._pack(q)
    assert torch.equal(qp2, cs.buf['qp'])
    if aux2 is not None:
        assert torch.equal(aux2, cs.buf['aux'])
    ks = env.split_keys((0, 1234), n + 1)
    assert (P.rng_bits(ks) == tf.split(tf.prng_key(1234), n + 1)).all()
    part = env.split_keys((0, 1234), n + 1, first=17, count=20)
    assert (P.rng_bits(part) == tf.split(tf.prng_key(1234), n + 1)[17:37]).all()


def test_shard_equivalence_and_invariants():
    """Envs never communicate: a run over N envs equals the concatenation of its shards bit for bit; at a
    BASELINE-sized batch the state keeps its invariants (unit quaternions, finite values, 0/1 flags)."""
    n, T = 1 << 16, 8
    keys = torch.as_tensor(P.keys_for(n, seed=11).view(np.int32))
    g = torch.Generator(device='cuda').manual_seed(0)
    acts = torch.rand((T, n, 8), device='cuda', generator=g) * 2 - 1
    full = _make('ant_heavenhell', n)
    s = full.reset(keys)
    for t in range(T):
        s = full.step(s, acts[t])
    halves = []
    for lo, hi in ((0, n // 2), (n // 2, n)):
        e = _make('ant_heavenhell', hi - lo)
        h = e.reset(keys[lo:hi])
        for t in range(T):
            h = e.step(h, acts[t, lo:hi].contiguous())
        halves.append(h)
    for name in ('qp', 'aux'):
        assert torch.equal(s.buf[name], torch.cat([h.buf[name] for h in halves], dim=1)), name
    for name in ('obs', 'reward', 'done', 'steps', 'rng'):
        assert torch.equal(s.buf[name], torch.cat([h.buf[name] for h in halves], dim=0)), name
    q = s.qp
    assert torch.isfinite(s.obs).all() and torch.isfinite(q.pos).all()
    assert (q.rot[:, :9].norm(dim=-1) - 1).abs().max() < 1e-5
    assert ((s.done == 0) | (s.done == 1)).all()
    assert torch.equal(q.rot[:, 9:, 0], torch.ones_like(q.rot[:, 9:, 0]))  # frozen bodies untouched


def test_errors_are_loud():
    from po_brax_b200 import envs
    with pytest.raises(ValueError):
        envs.create('ant', batch_size=-2)
    assert envs.create('ant', batch_size=0).unbatched     # __init__.py:64 `if batch_size:` -- 0 means un-vmapped too
    with pytest.raises(KeyError):
        envs.create('humanoid', batch_size=4)
    env = envs.create('ant_tag', batch_size=4)
    with pytest.raises(ValueError):
        env.reset(np.zeros((3, 2), np.uint32))
    s = env.reset(P.keys_for(4))
    with pytest.raises(ValueError):
        env.step(s, torch.zeros((4, 7), device='cuda'))
    with pytest.raises(RuntimeError):
        envs.create('ant_gather', batch_size=4, n_apples=20)


def test_gym_adapters_follow_the_reference_key_chain():
    """create_gym_env -> AutoresetVmapGymWrapper + EvalGymWrapper (scratch.py:17-23 usage): the env keys are
    split(PRNGKey(seed), N+1)[1:], the stored gym key advances to split(...)[0] on every reset."""
    from po_brax_b200 import envs
    n = 32
    e = envs.create_gym_env('ant_heavenhell', batch_size=n, seed=3, episode_length=5, eval_metrics=True, discount=0.99)
    obs = e.reset()
    oenv = oenvs.AntHeavenHellEnv()
    ks = tf.split(tf.prng_key(3), n + 1)
    want = oenv.reset(ks[1:])
    P.assert_qp_close(e.env._state.qp, want.qp, 'gym reset')
    assert tuple(obs.shape) == (n, 114)
    assert list(e.env._key) == ks[0].tolist()
    total_done = 0
    for t in range(11):
        a = torch.zeros((n, 8), device='cuda')
        o, r, d, info = e.step(a)
        total_done += int(d.sum())
        if t in (4, 9):  # episode_length 5 -> every env is done, reset from the advanced gym key
            assert bool(d.all()) and (e.env._state.info['steps'] == 0).all()
    st = e.get_stats()
    assert total_done >= 2 * n and abs(st['charts/mean_episodic_length'] - 5.0) < 1.0
    ks2 = tf.split(ks[0], n + 1)
    assert list(e.env._key) != ks[0].tolist()
    fresh = oenv.reset(ks2[1:])  # first re-reset used the keys drawn from ks[0]
    assert fresh.obs.shape == (n, 114)


@pytest.mark.parametrize('batch', [48, None])
def test_eval_gym_wrapper_statistics_equal_the_reference_recount(batch):
    """EvalGymWrapper (wrappers.py:175-229) through the fused pobrax_eval_update launch: the three means of get_stats()
    equal a host recount that follows the reference line by line (running sums, queues of finished episodes, nanmean)."""
    from po_brax_b200 import envs
    disc = 0.9
    e = envs.create_gym_env('ant_tag', batch_size=batch, seed=5, episode_length=6, eval_metrics=True, discount=disc)
    e.reset()
    n = batch or 1
    ret, dret, ln, cur = np.zeros(n), np.zeros(n), np.zeros(n, int), np.ones(n)
    rq, dq, lq = [], [], []
    g = torch.Generator(device='cuda').manual_seed(9)
    for t in range(20):
        a = torch.rand((n, 8) if batch else (8,), device='cuda', generator=g) * 2 - 1
        o, r, d, info = e.step(a)
        r = np.atleast_1d(np.asarray(r.cpu(), np.float64)); d = np.atleast_1d(np.asarray(d.cpu())) != 0
        ret += r; ln += 1; dret += r * cur; cur *= disc
        for i in np.nonzero(d)[0]:
            rq.append(ret[i]); dq.append(dret[i]); lq.append(ln[i])
        ret[d] = 0; dret[d] = 0; ln[d] = 0; cur[d] = 1
    st = e.get_stats()
    assert len(rq) >= 3 * n
    assert abs(st['charts/mean_episodic_return'] - np.mean(rq)) < 1e-5
    assert abs(st['charts/mean_discounted_episodic_return'] - np.mean(dq)) < 1e-5
    assert abs(st['charts/mean_episodic_length'] - np.mean(lq)) < 1e-9


@pytest.mark.parametrize('kind', ['ant_tag', 'ant_gather'])
def test_gym_autoreset_device_key_chain_equals_host_round_trip(kind):
    """AutoresetVmapGymWrapper.step (wrappers.py:245-262): with the gym key chain on the device
    (pobrax_reset_where_done_chain, no `done.any()` round trip) every buffer and the stored gym key are bit-identical
    to the reference-literal host path -- keys are drawn only on steps where some env finished."""
    from po_brax_b200 import envs
    n, T = 48, 40
    a_env = envs.create_gym_env(kind, batch_size=n, seed=5, episode_length=13)
    b_env = envs.create_gym_env(kind, batch_size=n, seed=5, episode_length=13)
    assert a_env.sync_free
    b_env.sync_free = False
    oa, ob = a_env.reset(), b_env.reset()
    assert torch.equal(oa, ob)
    g = torch.Generator(device='cuda').manual_seed(11)
    draws = 0
    for t in range(T):
        act = torch.rand((n, 8), device='cuda', generator=g) * 2 - 1
        key_before = b_env._key
        ra, rb = a_env.step(act), b_env.step(act)
        draws += b_env._key != key_before
        for x, y in zip(ra[:3], rb[:3]):
            assert torch.equal(x, y), t
        sa, sb = a_env._state, b_env._state
        for name in ('qp', 'aux', 'steps', 'truncation', 'rng'):
            if sa.buf.get(name) is not None:
                assert torch.equal(sa.buf[name], sb.buf[name]), (t, name)
    assert 0 < draws < T            # some steps finished an episode (keys drawn), some did not (gym key untouched)
    assert a_env._key == b_env._key  # pulled back from the device chain


@pytest.mark.parametrize('kind', ['ant_heavenhell', 'ant_tag'])
def test_randomized_autoreset_naive(kind):
    """wrappers.py:30-80: where done, qp/obs <- reset(info['rng']); Tag's key advances each step, HeavenHell's
    never does (every re-reset is the same state)."""
    from po_brax_b200 import envs
    from po_brax_b200.envs.wrappers import RandomizedAutoResetWrapperNaive
    n, L, T = 48, 3, 8
    keys = P.keys_for(n, seed=21)
    o = oenvs.RandomizedAutoResetNaive(oenvs.ENVS[kind](), episode_length=L)
    s = o.reset(keys)
    w = RandomizedAutoResetWrapperNaive(envs.create(kind, batch_size=n, episode_length=L, auto_reset=False))
    cs = w.reset(keys)
    rng = tf.prng_key(4)
    o.env.sys.track_margin = True
    dirty = np.zeros(n, bool)   # free-running: an env that passed a rounding-ambiguous branch stays out until it resets
    for t in range(T):
        rng, a = P.actions_for(rng, n)
        o.env.sys.margin = None
        s = o.step(s, a)
        cs = w.step(cs, torch.as_tensor(a, device='cuda'))
        dirty |= o.step_margin <= P.BRANCH_MARGIN
        dirty &= ~np.asarray(s.done, bool)   # done: qp is a fresh reset again
        assert np.array_equal(P.t2n(cs.info['steps']), s.info['steps']), t
        assert np.array_equal(P.t2n(cs.done)[~dirty], np.asarray(s.done, np.float32)[~dirty]), t
        assert (P.rng_bits(cs.info['rng']) == s.info['rng']).all()
        P.assert_qp_close(cs.qp, s.qp, f'{kind} randomized autoreset t={t}', vel_atol=5e-3, pos_scale=10.0, rows=~dirty)
        assert (~dirty).mean() >= 0.6   # up to L = 3 free-running steps of accumulated ambiguity
        if (t + 1) % L == 0:  # every env just hit the episode limit: qp is a fresh reset, bit-exact frozen bodies
            assert np.array_equal(P.t2n(cs.qp.pos)[:, 10:], s.qp.pos[:, 10:])


def test_randomized_autoreset_cached_refreshes_first_state():
    from po_brax_b200 import envs
    from po_brax_b200.envs.wrappers import RandomizedAutoResetWrapperCached
    n = 32
    keys = P.keys_for(n, seed=22)
    w = RandomizedAutoResetWrapperCached(envs.create('ant_tag', batch_size=n, episode_length=1000), n_steps_between_updates=3)
    cs = w.reset(keys)
    first0 = cs.buf['first_qp'].clone()
    a = torch.zeros((n, 8), device='cuda')
    cs = w.step(cs, a); cs = w.step(cs, a)
    assert torch.equal(cs.buf['first_qp'], first0)
    rng_before = P.rng_bits(cs.info['rng']).copy()
    cs = w.step(cs, a)  # 3rd call: refresh from reset(split(rng)[1]); rng <- split(rng)[0], then the step advances it
    assert not torch.equal(cs.buf['first_qp'], first0)
    want_rng = tf.split(tf.split(rng_before, 2)[:, 0], 2)[:, 0]
    assert (P.rng_bits(cs.info['rng']) == want_rng).all()
    fresh = oenvs.AntTagEnv().reset(tf.split(rng_before, 2)[:, 1])
    P.assert_qp_close(w.env._unpack(cs.buf['first_qp'], cs.buf['first_aux']), fresh.qp, 'refreshed first_qp')


def test_action_repeat_scales_dt_and_substeps():
    """ActionRepeatWrapper (wrappers.py:16-24): dt *= k, substeps *= k, i.e. 20 substeps of the same h."""
    n = 32
    keys = P.keys_for(n, seed=23)
    oenv = oenvs.AntHeavenHellEnv(action_repeat=2)
    assert oenv.sys.substeps == 20
    s = oenv.reset(keys)
    env = _make('ant_heavenhell', n, action_repeat=2, auto_reset=False)
    cs = env.reset(keys)
    rng = tf.prng_key(9)
    oenv.sys.track_margin = True
    dirty = np.zeros(n, bool)
    for t in range(3):
        rng, a = P.actions_for(rng, n)
        oenv.sys.margin = None
        s = oenv.step(s, a)
        cs = env.step(cs, torch.as_tensor(a, device='cuda'))
        dirty |= oenv.sys.margin <= P.BRANCH_MARGIN
        P.assert_qp_close(cs.qp, s.qp, f'action_repeat t={t}', vel_atol=5e-3, pos_scale=10.0, rows=~dirty)
    assert (~dirty).mean() > 0.25  # 60 substeps: many envs pass through a rounding-ambiguous branch


@pytest.mark.parametrize('kind,n', [('ant', 4096), ('ant_gather', 16384), ('ant_tag', 65536)])
def test_baseline_configs_invariants(kind, n):
    """BASELINE configs 2-4 at their full sizes with autoreset: size-independent properties over a 120-step
    rollout with a short episode limit (so every env resets several times)."""
    from po_brax_b200 import standard_observability_masks as M
    env = _make(kind, n, episode_length=25, auto_reset=True, eval_metrics=True)
    keys = env.split_keys((0, 77), n + 1, first=1, count=n)
    s = env.reset(keys)
    first_obs = s.info['first_obs'].clone()
    g = torch.Generator(device='cuda').manual_seed(3)
    done_total = 0
    for t in range(120):
        a = torch.rand((n, 8), device='cuda', generator=g) * 2 - 1
        s = env.step(s, a)
        d = s.done.bool()
        done_total += int(d.sum())
        assert torch.equal(s.obs[d], first_obs[d])                 # cached autoreset: obs <- first_obs where done
        assert (s.info['steps'] <= 25).all() and ((s.done == 0) | (s.done == 1)).all()
    q = s.qp
    assert torch.isfinite(s.obs).all() and torch.isfinite(q.pos).all() and torch.isfinite(q.vel).all()
    assert (q.rot[:, :9].norm(dim=-1) - 1).abs().max() < 1e-5    # unit quaternions
    assert torch.equal(q.vel[:, 9:], torch.zeros_like(q.vel[:, 9:]))  # frozen bodies never move by physics
    acc = dict(zip(env.ACC_NAMES, s.buf['acc'].tolist()))
    assert acc['episodes'] == done_total >= 4 * n and acc['sum_length'] <= 25 * acc['episodes']
    if kind == 'ant':   # standard_observability_masks: the three 'ant' index sets partition the 87 columns
        o = s.obs
        parts = [M.apply_mask(o, M.POSITION['ant']), M.apply_mask(o, M.VELOCITY['ant']), M.apply_mask(o, M.CFRC['ant'])]
        assert [p.shape[1] for p in parts] == [13, 14, 60] and torch.equal(torch.cat(parts, 1), o)
        assert (parts[2].abs() <= 1).all()                          # contact columns are clipped to [-1, 1]
    if kind == 'ant_gather':
        obj = q.pos[:, 11:]
        on_grid = (obj[..., :2] == obj[..., :2].round()).all(-1) & (obj[..., :2].abs() <= 6).all(-1)
        waiting = (obj == torch.tensor([18., 18., 12.], device='cuda')).all(-1)
        assert (on_grid | waiting).all()                             # objects sit on the integer grid or wait at (18,18,12)
        assert (s.obs[:, -20:] >= 0).all() and (s.obs[:, -20:] <= 1).all()
    if kind == 'ant_tag':
        tgt = q.pos[:, 10]
        assert (tgt[:, :2].abs() <= 4.5 + 1e-6).all()               # the opponent never leaves its cage


@pytest.mark.parametrize('mask', ['position', 'velocity', 'cfrc'])
def test_fused_observability_mask(mask):
    """standard_observability_masks 'ant' rows fused into the store: the masked env emits exactly obs[:, MASK]."""
    from po_brax_b200 import standard_observability_masks as M
    n = 200  # not a multiple of 8 or 32: exercises the ragged last warp
    keys = P.keys_for(n, seed=31)
    full, part = _make('ant', n, episode_length=4), _make('ant', n, episode_length=4, obs_mask=mask)
    idx = {'position': M.POSITION, 'velocity': M.VELOCITY, 'cfrc': M.CFRC}[mask]['ant']
    assert part.observation_size == len(idx)
    a, b = full.reset(keys), part.reset(keys)
    assert torch.equal(M.apply_mask(a.obs, idx), b.obs)
    g = torch.Generator(device='cuda').manual_seed(5)
    for t in range(9):
        act = torch.rand((n, 8), device='cuda', generator=g) * 2 - 1
        a, b = full.step(a, act), part.step(b, act)
        assert torch.equal(M.apply_mask(a.obs, idx), b.obs), t
        assert torch.equal(a.buf['qp'], b.buf['qp']) and torch.equal(a.reward, b.reward) and torch.equal(a.done, b.done)


def test_gather_column_range_and_cached_rows_through_the_out_of_line_obs_path():
    """Gather's step kernel sends the uncommon observation-row work (a column subset, rows rewritten from the cached first
    observation on autoreset) through the out-of-line write_obs_rows_general: same bits as the full-row kernel."""
    n, lo, hi = 77, 29, 150
    keys = P.keys_for(n, seed=13)
    full = _make('ant_gather', n, episode_length=3, auto_reset=True)
    part = _make('ant_gather', n, episode_length=3, auto_reset=True, obs_mask=(lo, hi))
    assert part.observation_size == hi - lo
    a, b = full.reset(keys), part.reset(keys)
    assert torch.equal(a.obs[:, lo:hi], b.obs)
    g = torch.Generator(device='cuda').manual_seed(6)
    for t in range(8):   # episode_length 3: every env restarts from its cached first state at t = 2 and t = 5
        act = torch.rand((n, 8), device='cuda', generator=g) * 2 - 1
        a, b = full.step(a, act), part.step(b, act)
        assert torch.equal(a.obs[:, lo:hi], b.obs), t
        assert torch.equal(a.buf['qp'], b.buf['qp']) and torch.equal(a.reward, b.reward) and torch.equal(a.done, b.done)
        if t in (2, 5):
            assert bool(a.done.all()) and torch.equal(a.obs, a.info['first_obs'])


def test_ragged_and_tiny_batches():
    """Batch sizes that do not fill a warp / CTA, and batch_size None (one env)."""
    for kind in KINDS:
        for n in (1, 5, 33):
            keys = P.keys_for(64, seed=41)[:n]
            big = _make(kind, 64)
            small = _make(kind, n)
            sb, ss = big.reset(P.keys_for(64, seed=41)), small.reset(keys)
            a = torch.zeros((64, 8), device='cuda')
            a[:, ::2] = 0.3
            for _ in range(3):
                sb, ss = big.step(sb, a), small.step(ss, a[:n].contiguous())
            assert torch.equal(sb.obs[:n], ss.obs) and torch.equal(sb.buf['qp'][:, :n], ss.buf['qp'])


def test_unvmapped_create_has_no_batch_axis():
    """create(batch_size=None) adds no VmapWrapper (__init__.py:64): reset takes one key, State fields carry no batch
    axis -- obs (D,), scalar reward / done, qp.pos (nb, 3) -- and equal env 0 of a batch built from the same key."""
    from po_brax_b200 import envs
    for kind in KINDS:
        one = envs.create(kind, auto_reset=False)
        key = P.keys_for(1)[0]
        s = one.reset(key)
        ref_env = _make(kind, 1, auto_reset=False)
        r = ref_env.reset(key[None])
        D, nb = one.observation_size, one.num_bodies
        assert tuple(s.obs.shape) == (D,) and s.reward.dim() == 0 and s.done.dim() == 0
        assert tuple(s.qp.pos.shape) == (nb, 3) and tuple(s.qp.rot.shape) == (nb, 4)
        assert all(v.dim() == 0 for v in s.metrics.values()) and s.info['steps'].dim() == 0
        if kind != 'ant':
            assert tuple(s.info['rng'].shape) == (2,)
        a = torch.full((8,), 0.25, device='cuda')
        s, r = one.step(s, a), ref_env.step(r, a[None])
        assert torch.equal(s.obs, r.obs[0]) and torch.equal(s.reward, r.reward[0]) and torch.equal(s.qp.pos, r.qp.pos[0])
        s2 = s.replace(obs=torch.zeros(D, device='cuda'))
        assert tuple(s2.obs.shape) == (D,)


def test_eval_metrics_accumulators_match_a_host_recount():
    """create(..., eval_metrics=True): the device-side accumulators equal what a host loop counts from done / reward."""
    n, L = 512, 7
    env = _make('ant_heavenhell', n, episode_length=L, auto_reset=True, eval_metrics=True)
    s = env.reset(P.keys_for(n, seed=51))
    g = torch.Generator(device='cuda').manual_seed(7)
    ret = torch.zeros(n, device='cuda', dtype=torch.float64)
    episodes = sum_ret = sum_len = trunc = 0.0
    for t in range(30):
        s = env.step(s, torch.rand((n, 8), device='cuda', generator=g) * 2 - 1)
        ret += s.reward.double()
        d = s.done.bool()
        episodes += float(d.sum()); sum_ret += float(ret[d].sum()); sum_len += float(s.info['steps'][d].sum())
        trunc += float(s.info['truncation'].sum())
        ret[d] = 0
    em = {k: float(v) for k, v in s.info['eval_metrics'].items()}
    assert em['episodes'] == episodes and em['sum_length'] == sum_len and em['truncations'] == trunc
    assert abs(em['sum_return'] - sum_ret) < 1e-6


@pytest.mark.parametrize('kind,kw', [
    ('ant_gather', dict(n_apples=3, n_bombs=2, n_bins=6, cage_xy=(4, 5), catch_range=1.5, sensor_range=4.0, dying_cost=-3.0)),
    ('ant_tag', dict(tag_radius=2.0, visible_radius=6.0, target_step=0.25, min_spawn_distance=3.0, cage_xy=(3.5, 4.0), dying_cost=-0.5)),
    ('ant_heavenhell', dict(heaven_hell=((-4.0, 6.0), (4.0, 6.0)), priest_position=(0.0, 6.0), visible_radius=2.5, dying_cost=-1.0)),
])
def test_constructor_kwargs_of_the_reference_envs(kind, kw):
    """Non-default constructor arguments (ant_gather.py:59-69, ant_tag.py:38-45, ant_heavenhell.py:51-56) reach the
    kernels: reset parity + teacher-forced steps against the oracle built with the same arguments."""
    n, T = 96, 6
    keys = P.keys_for(n, seed=61)
    oenv = oenvs.ENVS[kind](**kw)
    env = _make(kind, n, auto_reset=False, **kw)
    nb = oenv.sys.num_bodies
    assert env.observation_size == oenv.reset(keys[:2]).obs.shape[1] and env.num_bodies == nb
    s = oenv.reset(keys)
    cs = env.reset(keys)
    assert np.array_equal(P.t2n(cs.qp.pos)[:, 9:], s.qp.pos[:, 9:])
    P.assert_qp_close(cs.qp, s.qp, f'{kind} kwargs reset')
    rng = tf.prng_key(8)
    oenv.sys.track_margin = True
    for t in range(T):
        rng, a = P.actions_for(rng, n)
        c0 = env.state_from_qp(P.qp_to_torch(s.qp), rng=s.info.get('rng'))
        oenv.sys.margin = None
        nxt = oenv.step(oenvs.State(s.qp.copy(), s.obs, s.reward, s.done, dict(s.metrics), dict(s.info)), a)
        got = env.step(c0, torch.as_tensor(a, device='cuda'))
        clear = oenv.sys.margin > P.BRANCH_MARGIN
        P.assert_qp_close(got.qp, nxt.qp, f'{kind} kwargs t={t}', rows=clear)
        assert np.array_equal(P.t2n(got.done), np.asarray(nxt.done, np.float32))
        assert np.array_equal(P.t2n(got.reward), nxt.reward)
        mask = np.ones_like(nxt.obs, bool)
        mask[~clear] = False
        if kind == 'ant_gather':
            mask[:, -2 * oenv.n_bins:] &= _gather_reading_mask(oenv, nxt, got)
        P.assert_obs_close(P.t2n(got.obs), nxt.obs, kind, nb, f'{kind} kwargs t={t}', mask=mask)
        s = nxt


def test_unbatched_gym_env_follows_the_brax_gym_key_chain():
    """create_gym_env(batch_size=None) -> AutoresetGymWrapper (wrappers.py:232-237 over brax GymWrapper): reset draws
    key1, key2 = split(key), resets from key2 and keeps key1; a finished episode triggers a FULL reset from the chain
    (fresh info['rng']) whose observation is returned with the finished step's reward / done."""
    from po_brax_b200 import envs
    e = envs.create_gym_env('ant_tag', seed=7, episode_length=4)
    oenv = oenvs.AntTagEnv()
    obs = e.reset()
    ks = tf.split(tf.prng_key(7), 2)
    want = oenv.reset(ks[1:2])
    assert tuple(obs.shape) == (103,) and e.observation_space.shape == (103,) and e.action_space.shape == (8,)
    cols = list(range(29)) + [101, 102]   # everything but the contact columns (resting-foot sign ambiguity)
    assert np.allclose(P.t2n(obs)[cols], want.obs[0, cols], atol=2e-5)
    assert list(e._key) == ks[0].tolist()
    for t in range(4):
        o, r, d, info = e.step(np.zeros(8, np.float32))
        assert o.shape == (103,) and r.dim() == 0 and d.dim() == 0 and set(info) == {'hits'}
        assert bool(d) == (t == 3)            # truncated by episode_length 4 (no tag, no death from rest)
    ks2 = tf.split(ks[0], 2)
    want2 = oenv.reset(ks2[1:2])
    assert np.allclose(P.t2n(o)[cols], want2.obs[0, cols], atol=2e-5)
    assert (P.rng_bits(e._state.info['rng']) == want2.info['rng'][0]).all()   # full reset: fresh info['rng']
    assert list(e._key) == ks2[0].tolist()
    assert float(e._state.info['steps']) == 0.0 and float(e._state.done) == 0.0
    with pytest.raises(ValueError):
        envs.create_gym_env('ant_tag', batch_size=0)


@pytest.mark.parametrize('kind', ['ant_tag', 'ant_heavenhell'])
def test_gym_step_as_cuda_graph_equals_plain_launches(kind):
    """create_gym_env(..., cuda_graph=True) replays (step + device-side autoreset) as one CUDA graph: every output of
    every step, the State and the gym key chain must equal the plain launches bit for bit, across resets (a reset
    re-captures) and with actions given as host arrays."""
    from po_brax_b200 import envs
    n = 300
    g = torch.Generator(device='cuda').manual_seed(11)
    acts = torch.rand((45, n, 8), device='cuda', generator=g) * 2 - 1
    outs = []
    for graph in (False, True):
        e = envs.create_gym_env(kind, batch_size=n, seed=5, episode_length=7, cuda_graph=graph)
        rec = [e.reset().clone()]
        for t in range(45):
            if t == 20:
                rec.append(e.reset().clone())           # new State buffers: the graph path must re-capture
            a = acts[t] if t % 2 else acts[t].cpu().numpy()
            o, r, d, info = e.step(a)
            rec += [o.clone(), r.clone(), d.clone()] + [v.clone() for v in info.values()]
        rec += [v.clone() for v in e._state.buf.values() if v is not None]
        assert (e._graph is not None) == graph
        rec.append(torch.tensor(list(e._key)))          # pulls the key chain back to the host (drops the graph)
        torch.cuda.synchronize()
        outs.append(rec)
    assert len(outs[0]) == len(outs[1])
    for i, (a, b) in enumerate(zip(*outs)):
        assert torch.equal(a, b), (kind, i)


def test_gym_outputs_of_step_t_survive_step_t_plus_1():
    """The reference returns fresh arrays every step; rollout code keeps them (`dones.append(d)`). The gym adapters
    return copies by default; copy=False hands out the live buffers (valid until the next step)."""
    from po_brax_b200 import envs
    n = 64
    for graph in (False, True):
        e = envs.create_gym_env('ant_tag', batch_size=n, seed=3, episode_length=3, cuda_graph=graph)
        obs0 = e.reset()
        keep0 = obs0.clone()
        g = torch.Generator(device='cuda').manual_seed(5)
        held, snaps = [], []
        for t in range(5):
            out = e.step(torch.rand((n, 8), device='cuda', generator=g) * 2 - 1)
            held.append(out)
            snaps.append((out[0].clone(), out[1].clone(), out[2].clone(), {k: v.clone() for k, v in out[3].items()}))
        assert torch.equal(obs0, keep0)
        for (o, r, d, m), (so, sr, sd, sm) in zip(held, snaps):
            assert torch.equal(o, so) and torch.equal(r, sr) and torch.equal(d, sd)
            assert all(torch.equal(m[k], sm[k]) for k in m)
        assert any(bool(d.any()) for _, _, d, _ in held) and not bool(held[0][2].all())   # dones differ over the steps
    live = envs.create_gym_env('ant_tag', batch_size=n, seed=3, episode_length=3, copy=False)
    live.reset()
    o1 = live.step(torch.zeros((n, 8), device='cuda'))[0]
    assert o1.data_ptr() == live._state.buf['obs'].data_ptr()


def test_tag_done_is_bool_after_a_step():
    """ant_tag.py:88 vs :127: done is f32 zeros at reset and `jp.logical_or(dead, hit)` (bool) after a step; the
    other envs stay f32."""
    for kind in KINDS:
        env = _make(kind, 16, auto_reset=False)
        s = env.reset(P.keys_for(16))
        assert s.done.dtype == torch.float32
        s = env.step(s, torch.zeros((16, 8), device='cuda'))
        assert s.done.dtype == (torch.bool if kind == 'ant_tag' else torch.float32)
        assert torch.equal(s.done.float(), s.buf['done'])


def test_env_sys_surface():
    """What code outside the envs reads from env.sys (wrappers.py:22-23,140; ant_heavenhell.py:64-69,90)."""
    from oracle import config as ocfg
    for kind, cfg in (('ant', ocfg.ant_config()), ('ant_heavenhell', None), ('ant_tag', None), ('ant_gather', None)):
        env = _make(kind, 4, action_repeat=2)
        o = oenvs.ENVS[kind](action_repeat=2)
        assert env.sys.num_bodies == o.sys.num_bodies == env.num_bodies
        assert env.sys.body.index == o.sys.index
        assert env.sys.num_joint_dof == 8 and env.action_size == env.sys.num_joint_dof + env.sys.num_forces_dof
        assert abs(env.sys.config.dt - 0.1) < 1e-7 and env.sys.config.substeps == 20
        assert np.allclose(P.t2n(env.sys.default_angle()), o.sys.default_angle(), atol=1e-7)
        assert env.torso_idx == o.torso_idx
        for name in ('target_idx', 'hell_idx', 'priest_idx'):
            if hasattr(o, name):
                assert getattr(env, name) == getattr(o, name)
        with pytest.raises(AttributeError):
            env.sys.step


def test_eval_metrics_record_has_the_brax_shape():
    """create(eval_metrics=True) -> state.info['eval_metrics'] is an EvalMetrics record (brax EvalWrapper): per-env
    current return, completed sums, episode count and total steps."""
    n, L = 256, 6
    env = _make('ant_tag', n, episode_length=L, auto_reset=True, eval_metrics=True)
    s = env.reset(P.keys_for(n, seed=2))
    g = torch.Generator(device='cuda').manual_seed(9)
    cur = torch.zeros(n, device='cuda')
    done_sum = episodes = steps = hits = 0.0
    for t in range(20):
        s = env.step(s, torch.rand((n, 8), device='cuda', generator=g) * 2 - 1)
        cur += s.reward
        d = s.done
        done_sum += float(cur[d].double().sum()); episodes += float(d.sum()); steps += float(s.info['steps'][d].sum())
        hits += float(s.metrics['hits'].sum())
        cur[d] = 0
    em = s.info['eval_metrics']
    assert tuple(em.current_episode_metrics['reward'].shape) == (n,)
    assert torch.allclose(em.current_episode_metrics['reward'], cur, atol=1e-6)
    assert float(em.completed_episodes) == episodes and float(em.completed_episodes_steps) == steps
    assert abs(float(em.completed_episodes_metrics['reward']) - done_sum) < 1e-6
    assert float(em.completed_episodes_metrics['hits']) == hits
    assert em['episodes'] is em.acc['episodes']


def test_heavenhell_pack_refuses_unrepresentable_goal_rows():
    """State.replace(qp=...) / state_from_qp: the packed state keeps one goal-side flag, so Target / Hell / Priest
    rows other than the configured positions raise instead of snapping silently."""
    env = _make('ant_heavenhell', 8, auto_reset=False)
    s = env.reset(P.keys_for(8))
    q = s.qp
    s2 = s.replace(qp=q)                                   # the unpacked qp round-trips
    assert torch.equal(s2.buf['aux'], s.buf['aux'])
    swapped = q.pos.clone()
    swapped[:, [env.target_idx, env.hell_idx]] = swapped[:, [env.hell_idx, env.target_idx]]
    s3 = s.replace(qp=q.replace(pos=swapped))              # the other side is representable
    assert torch.equal(s3.buf['aux'][2], 1.0 - s.buf['aux'][2])
    bad = q.pos.clone()
    bad[3, env.target_idx, 0] += 1e-3
    with pytest.raises(ValueError):
        s.replace(qp=q.replace(pos=bad))
    bad = q.pos.clone()
    bad[0, env.priest_idx, 1] = 6.5
    with pytest.raises(ValueError):
        env.state_from_qp(q.replace(pos=bad))


def test_gather_fractional_cage_keeps_the_reference_grid():
    """ant_gather.py:88: arange(-cage, cage + 1) keeps a fractional origin (cage 4.5 -> -4.5 ... 4.5); object
    placement, the waiting area and the walls must match the oracle built with the same cage."""
    n = 128
    kw = dict(cage_xy=(4.5, 3.5), n_apples=4, n_bombs=3)
    oenv = oenvs.ENVS['ant_gather'](**kw)
    env = _make('ant_gather', n, auto_reset=False, **kw)
    keys = P.keys_for(n, seed=13)
    s, cs = oenv.reset(keys), env.reset(keys)
    assert np.array_equal(P.t2n(cs.qp.pos)[:, oenv.obj], s.qp.pos[:, oenv.obj])
    assert (np.abs(s.qp.pos[:, oenv.obj, 0]) % 1 == 0.5).all()
    # walk the objects into the waiting area: done-if-all-collected compares against the same waiting position
    far = s.qp.pos.copy()
    far[:, oenv.obj] = oenv.waiting_area
    far[0, oenv.obj[0]] = s.qp.pos[0, oenv.obj[0]]
    from oracle import brax_v1 as bx
    qp = bx.QP(far, s.qp.rot, s.qp.vel, s.qp.ang)
    nxt = oenv.step(oenvs.State(qp.copy(), s.obs, s.reward, s.done, dict(s.metrics), dict(s.info)), np.zeros((n, 8), np.float32))
    got = env.step(env.state_from_qp(P.qp_to_torch(qp), rng=s.info['rng']), torch.zeros((n, 8), device='cuda'))
    assert np.array_equal(P.t2n(got.done), nxt.done) and nxt.done[1:].all() and not nxt.done[0]


def test_raw_c_abi_without_the_facade():
    """pobrax_create / reset / step / unpack_qp / destroy driven straight through ctypes on caller-owned device
    buffers (torch only allocates them): what a reference-side binding would do."""
    import ctypes as C
    from po_brax_b200 import _lib
    lib = C.CDLL(_lib.LIB_PATH)
    n = 96
    p = _lib.PobraxParams()
    assert lib.pobrax_default_params(_lib.ANT_HEAVENHELL, C.byref(p)) == 0
    p.num_envs, p.auto_reset = n, _lib.AUTORESET_OFF
    L = _lib.PobraxLayout()
    assert lib.pobrax_layout(C.byref(p), C.byref(L)) == 0
    h = C.c_void_p()
    assert lib.pobrax_create(C.byref(p), 0, C.byref(h)) == 0
    f = dict(dtype=torch.float32, device='cuda:0')
    buf = {'qp': torch.zeros((L.qp_planes, n, 4), **f), 'aux': torch.zeros((L.aux_dim, n), **f),
           'obs': torch.zeros((n, L.obs_dim), **f), 'reward': torch.zeros(n, **f), 'done': torch.zeros(n, **f),
           'steps': torch.zeros(n, **f), 'truncation': torch.zeros(n, **f),
           'rng': torch.zeros((n, 2), dtype=torch.int32, device='cuda:0'), 'metrics': torch.zeros((L.metrics_dim, n), **f)}
    st = _lib.PobraxState()
    for k, v in buf.items():
        setattr(st, k, v.data_ptr())
    keys_np = P.keys_for(n, seed=0)
    keys = torch.from_numpy(keys_np.view(np.int32).copy()).cuda()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.pobrax_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pobrax_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.pobrax_unpack_qp.argtypes = [C.c_void_p] * 8
    lib.pobrax_last_error.restype = C.c_char_p
    assert lib.pobrax_reset(h, keys.data_ptr(), C.byref(st), stream) == 0, lib.pobrax_last_error()
    a_np = tf.uniform(tf.prng_key(1), n * 8, -1.0, 1.0).reshape(n, 8)
    act = torch.from_numpy(a_np).cuda()
    assert lib.pobrax_step(h, C.byref(st), act.data_ptr(), stream) == 0, lib.pobrax_last_error()
    nb = L.num_bodies
    pos, rot = torch.empty((n, nb, 3), **f), torch.empty((n, nb, 4), **f)
    vel, ang = torch.empty((n, nb, 3), **f), torch.empty((n, nb, 3), **f)
    assert lib.pobrax_unpack_qp(h, buf['qp'].data_ptr(), buf['aux'].data_ptr(), pos.data_ptr(), rot.data_ptr(),
                                vel.data_ptr(), ang.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    oenv = oenvs.AntHeavenHellEnv()
    o = oenv.step(oenv.reset(keys_np), a_np)
    assert (buf['rng'].cpu().numpy().view(np.uint32) == o.info['rng']).all()
    assert np.array_equal(buf['done'].cpu().numpy(), o.done) and np.array_equal(buf['reward'].cpu().numpy(), o.reward)
    assert np.abs(pos.cpu().numpy() - o.qp.pos).max() < 1e-4 and np.abs(rot.cpu().numpy() - o.qp.rot).max() < 1e-4
    assert np.array_equal(pos.cpu().numpy()[:, 9:], o.qp.pos[:, 9:])
    # errors come back as codes + a message, never as a crash
    assert lib.pobrax_step(h, C.byref(st), None, stream) != 0 and b'null' in lib.pobrax_last_error()
    assert lib.pobrax_destroy(h) == 0


def test_two_handles_on_two_devices():
    """One process driving two GPUs: kernel attributes / occupancy are set up per handle on its own device
    (Gather's reset needs > 48 KB of dynamic shared memory from cage_xy 7 on)."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    from po_brax_b200 import envs
    n, kw = 256, dict(cage_xy=(8, 8))
    outs = []
    for dev in (0, 1):
        env = envs.Env('ant_gather', batch_size=n, device=f'cuda:{dev}', **kw)
        keys = P.keys_for(n, seed=4)
        s = env.reset(keys)
        a = torch.full((n, 8), 0.1, device=f'cuda:{dev}')
        for _ in range(3):
            s = env.step(s, a)
        torch.cuda.synchronize(dev)
        outs.append((s.obs.cpu(), s.buf['qp'].cpu()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_small_batch_instantiation_is_bit_identical(monkeypatch):
    """Batches up to 2 warps per sub-partition run `step_kernel<KIND, true>` (the same source compiled for up to 255
    registers: one warp's latency through the substeps drops ~15 %). Results must not depend on which instantiation
    ran -- a shard must equal its slice of a larger batch -- so the two are held bit for bit against each other on all
    four families over 80 steps with short episodes (cached autoreset), HeavenHell's spawn-wall contacts included.
    POBRAX_SMALL_BATCH_ENVS is read at create."""
    from po_brax_b200 import envs
    n, T = 2048, 80
    g = torch.Generator(device='cuda').manual_seed(3)
    acts = torch.rand((4, n, 8), device='cuda', generator=g) * 2 - 1
    for kind in KINDS:
        outs = []
        for limit in ('0', '100000000'):          # never / always the small-batch instantiation
            monkeypatch.setenv('POBRAX_SMALL_BATCH_ENVS', limit)
            env = envs.create(kind, batch_size=n, episode_length=25)
            s = env.reset(P.keys_for(n, seed=17))
            for t in range(T):
                s = env.step(s, acts[t % 4])
            torch.cuda.synchronize()
            outs.append({k: s.buf[k].clone() for k in ('obs', 'qp', 'reward', 'done', 'steps')})
        for k in outs[0]:
            assert torch.equal(outs[0][k], outs[1][k]), (kind, k)
