"""Shared helpers for the GPU parity tests: oracle <-> CUDA state conversion and tolerances.

Tolerances (stated here once, used by every parity test):
  * integers / RNG bits / done flags / item indices: bit-exact.
  * pos, rot, joint angles (teacher-forced, one env step from identical states): |d| <= 1e-6 + 2e-6*|x|
    -- SURVEY App. C gate; the float32 oracle itself is 8e-7 (pos) / 4e-7 (rot) away from its float64 twin.
  * envs where a contact / actuator decision of the reference algorithm is rounding-ambiguous in this step (within
    BRANCH_MARGIN of its switching point; reported by the oracle itself, oracle/brax_v1.py:_note_margin) get the
    TWO-BRANCH check (class TwoBranch): the result must meet the same tight gates against the oracle's step with
    one of the ambiguous decisions taken the other way (oracle/brax_step_impl.h BranchCtl). Only envs that match
    none of those are "unexplained"; they are held to the loose bound and must stay under 1 % of the batch.
  * vel, ang, joint velocities, contact impulses: |d| <= 3e-4 -- the stiff joint springs (k = 18000, h = 5 ms)
    amplify float32 rounding of positions: the float32 oracle is up to 1.1e-4 away from the float64 oracle
    after one step (measured over 60 steps x 256 envs), so 3e-4 is ~3x the reference arithmetic's own noise.
"""
import numpy as np
import torch

from oracle import brax_v1 as bx
from oracle import envs as oenvs
from oracle import threefry as tf

POS_TOL = (1e-6, 2e-6)
VEL_ATOL = 3e-4
# An env counts as rounding-ambiguous when a contact is within this margin of a discontinuous branch
# (margins are in velocity units; penetration margins are scaled x100, see below).
BRANCH_MARGIN = 5e-4
LOOSE_POS, LOOSE_VEL = 2e-2, 3.0   # unexplained envs: one contact impulse / one actuator cut-off (350*h = 1.75 rad/s) apart
MAX_UNEXPLAINED = 0.01             # share of a batch allowed to match neither branch of its ambiguous decisions


def keys_for(n, seed=0):
    """VmapGymWrapper._reset key scheme (wrappers.py:160-163): split(PRNGKey(seed), n+1)[1:]."""
    return tf.split(tf.prng_key(seed), n + 1)[1:]


def actions_for(rng, n):
    """BASELINE config 1: rng, k = split(rng); a = uniform(k, (n, 8), -1, 1)."""
    ks = tf.split(rng, 2)
    return ks[0], tf.uniform(ks[1], n * 8, -1.0, 1.0).reshape(n, 8)


def t2n(x):
    return x.detach().cpu().numpy()


def rng_bits(t):
    return t2n(t).view(np.uint32)


def qp_to_torch(qp, device='cuda'):
    from po_brax_b200.envs import QP
    return QP(*[torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=device)
                for a in (qp.pos, qp.rot, qp.vel, qp.ang)])


def close_pos(a, b):
    return np.abs(a - b) <= POS_TOL[0] + POS_TOL[1] * np.abs(b)


def assert_qp_close(got, want, what='', vel_atol=VEL_ATOL, rows=None, loose=False, pos_scale=1.0):
    for name, tight in (('pos', True), ('rot', True), ('vel', False), ('ang', False)):
        g, w = t2n(getattr(got, name)), getattr(want, name)
        if rows is not None:
            g, w = g[rows], w[rows]
        if g.size == 0:
            continue
        if loose:
            ok = np.abs(g - w) <= (LOOSE_POS if tight else LOOSE_VEL)
        else:
            ok = (np.abs(g - w) <= pos_scale * (POS_TOL[0] + POS_TOL[1] * np.abs(w))) if tight else \
                 (np.abs(g - w) <= vel_atol)
        assert ok.all(), f'{what} qp.{name}: max |d| = {np.abs(g - w).max():.3e} at {np.argwhere(~ok)[:4].tolist()}'


def obs_tolerances(kind, nb, obs_dim):
    """Per-column absolute / relative tolerance of the observation."""
    P = 1 if kind == 'ant' else 3
    atol = np.full(obs_dim, VEL_ATOL, np.float64)
    rtol = np.zeros(obs_dim, np.float64)
    tight = list(range(0, P + 4))                 # torso pos, rot
    atol[tight] = POS_TOL[0]
    rtol[tight] = POS_TOL[1]
    atol[P + 4:P + 12] = 5e-6                     # joint angles (atan2 of unit-vector products)
    atol[P + 18:P + 26] = 2 * VEL_ATOL            # joint velocities: (w_parent - w_child).axis, two angular velocities
    end = P + 26 + 6 * nb
    atol[end:] = 1e-5                             # task extras (exact unless stated otherwise by the test)
    return atol, rtol


def assert_obs_close(got, want, kind, nb, what='', mask=None):
    atol, rtol = obs_tolerances(kind, nb, want.shape[-1])
    ok = np.abs(got - want) <= atol + rtol * np.abs(want)
    if mask is not None:
        ok = ok | ~mask
    assert ok.all(), (f'{what} obs: {int((~ok).sum())} entries off, worst {np.abs(got - want)[~ok].max():.3e} at '
                      f'{np.argwhere(~ok)[:6].tolist()}')


# ---------------------------------------------------------------------------------- two-branch check
class Got:
    """NumPy snapshot of an implementation's result for the 9 ant bodies: pos/rot/vel/ang [N,9,*] and the clipped
    Info.contact sums [N,9,3] x 2 (what the observation carries)."""

    def __init__(self, pos, rot, vel, ang, ccv, cca):
        self.pos, self.rot, self.vel, self.ang, self.ccv, self.cca = pos, rot, vel, ang, ccv, cca

    @classmethod
    def from_state(cls, state, kind, nb):
        q = state.qp
        obs = t2n(state.obs)
        c0 = (1 if kind == 'ant' else 3) + 26
        n = obs.shape[0]
        ccv = obs[:, c0:c0 + 3 * nb].reshape(n, nb, 3)[:, :9]
        cca = obs[:, c0 + 3 * nb:c0 + 6 * nb].reshape(n, nb, 3)[:, :9]
        return cls(t2n(q.pos)[:, :9], t2n(q.rot)[:, :9], t2n(q.vel)[:, :9], t2n(q.ang)[:, :9], ccv, cca)

    def take(self, idx):
        return Got(*[getattr(self, k)[idx] for k in ('pos', 'rot', 'vel', 'ang', 'ccv', 'cca')])


def physics_match(got: Got, qp1, info, contact_clipped=True):
    """bool[N]: `got` meets the TIGHT gates against the oracle result (qp1, info) -- pos / rot 1e-6 + 2e-6|x|,
    vel / ang / contact impulses 3e-4 -- on every ant body of the env."""
    ok = np.ones(got.pos.shape[0], bool)
    for g, w in ((got.pos, qp1.pos[:, :9]), (got.rot, qp1.rot[:, :9])):
        ok &= (np.abs(g - w) <= POS_TOL[0] + POS_TOL[1] * np.abs(w)).reshape(len(ok), -1).all(1)
    for g, w in ((got.vel, qp1.vel[:, :9]), (got.ang, qp1.ang[:, :9])):
        ok &= (np.abs(g - w) <= VEL_ATOL).reshape(len(ok), -1).all(1)
    clip = (lambda x: np.clip(x, -1, 1)) if contact_clipped else (lambda x: x)
    for g, w in ((got.ccv, info.contact_vel[:, :9]), (got.cca, info.contact_ang[:, :9])):
        ok &= (np.abs(g - clip(w)) <= VEL_ATOL).reshape(len(ok), -1).all(1)
    return ok


class TwoBranch:
    """For envs whose step had rounding-ambiguous decisions: does the result match the oracle with some of those
    decisions taken the other way? Uses the C twin of the oracle (tests/test_oracle_c.py pins it to the NumPy text):
    its k-th ambiguous predicate evaluation of an env is inverted where bit k of the env's flip mask is set."""

    def __init__(self, system_factory, threads=None, max_bits=4):
        import os
        from oracle import cstep
        self.sys = system_factory()
        self.be = cstep.attach(self.sys, threads=threads or os.cpu_count() or 1)
        self.max_bits = max_bits

    def explain(self, qp0, act, idx, got: Got, contact_clipped=True):
        """idx: env indices to explain. Returns (explained bool[len(idx)], n_marginal int[len(idx)])."""
        idx = np.asarray(idx)
        if idx.size == 0:
            return np.zeros(0, bool), np.zeros(0, np.int32)
        sub, a = qp0.take(idx), np.ascontiguousarray(act[idx])
        qp1, info = self.be.step(sub, a, flip_thr=BRANCH_MARGIN)
        nm = self.be.n_marginal.copy()
        ok = physics_match(got.take(idx), qp1, info, contact_clipped)
        for mask in range(1, 1 << self.max_bits):
            todo = ~ok & (nm >= mask.bit_length())
            if not todo.any():
                if not (~ok & (nm > mask.bit_length())).any():
                    break
                continue
            t = np.nonzero(todo)[0]
            qp1, info = self.be.step(sub.take(t), a[t], flip_mask=np.full(len(t), mask, np.uint32),
                                     flip_thr=BRANCH_MARGIN)
            ok[t] |= physics_match(got.take(idx[t]), qp1, info, contact_clipped)
        return ok, nm
