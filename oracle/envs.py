"""CPU restatement (NumPy, batched over envs = the reference's jax.vmap axis) of the po-brax task
envs and of the wrappers `po_brax.envs.create` stacks on them.

TEST INFRASTRUCTURE ONLY (see oracle/threefry.py header).

Follows, function by function:
  AntHeavenHellEnv  /root/reference/po_brax/envs/ant_heavenhell.py:75-158
  AntGatherEnv      /root/reference/po_brax/envs/ant_gather.py:93-213
  AntTagEnv         /root/reference/po_brax/envs/ant_tag.py:63-181
  create()          /root/reference/po_brax/envs/__init__.py:50-72   (Episode -> Vmap -> AutoReset)
  gym autoreset     /root/reference/po_brax/envs/wrappers.py:160-166,245-262
  plain Ant, EpisodeWrapper, AutoResetWrapper: brax v0.0.12 (brax/envs/ant.py, brax/envs/wrappers.py),
    un-vendored; restated from the published source.
Parity status: RNG-driven integer outcomes are pinned by the threefry KATs; the task logic has no
reference artefact (the reference has no tests) => "parity unpinned" beyond the physics fixture.
"""
import numpy as np

from . import brax_v1 as bx
from . import config as cfgs
from . import threefry as tf

F = np.float32


class State:
    """brax.envs.env.State (batched): qp, obs[N,D], reward[N], done[N], metrics{}, info{}."""

    def __init__(self, qp, obs, reward, done, metrics, info):
        self.qp, self.obs, self.reward, self.done, self.metrics, self.info = qp, obs, reward, done, metrics, info

    def replace(self, **kw):
        d = dict(qp=self.qp, obs=self.obs, reward=self.reward, done=self.done,
                 metrics=self.metrics, info=self.info)
        d.update(kw)
        return State(**d)


def _clip1(x):
    return np.clip(x, F(-1), F(1))


def _norm2(v):
    return np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1])


class _AntBase:
    """Shared pieces: joint sampling, dead test, the common obs layout."""
    action_size = 8

    def __init__(self, cfg, action_repeat=1, walls=True, dtype=np.float32):
        if action_repeat != 1:
            cfg = cfgs.scale_action_repeat(cfg, action_repeat)
        self.sys = bx.System(cfg, dtype=dtype, walls=walls)
        self.torso_idx = self.sys.index['$ Torso']

    def _sample_joints(self, k_pos, k_vel):
        qpos = self.sys.default_angle() + tf.uniform(k_pos, 8, -.1, .1)
        qvel = tf.uniform(k_vel, 8, -.1, .1)
        return qpos.astype(F), qvel.astype(F)

    def _dead(self, qp):
        z = qp.pos[:, self.torso_idx, 2]
        dead = np.where(z < F(0.2), F(1), F(0))
        return np.where(z > F(1.0), F(1), dead).astype(F)

    def _common_obs(self, qp, info, full_pos=True):
        ja, jv = self.sys.angle_vel(qp)
        N = qp.pos.shape[0]
        p0 = qp.pos[:, 0] if full_pos else qp.pos[:, 0, 2:]
        return [p0, qp.rot[:, 0], ja, qp.vel[:, 0], qp.ang[:, 0], jv,
                _clip1(info.contact_vel).reshape(N, -1), _clip1(info.contact_ang).reshape(N, -1)]

    @property
    def observation_size(self):
        return self.reset(tf.prng_key(0)[None]).obs.shape[-1]


class AntEnv(_AntBase):
    """brax.envs.ant.Ant (v0.0.12), registered at /root/reference/po_brax/envs/__init__.py:30."""

    def __init__(self, **kw):
        super().__init__(cfgs.ant_config(), **kw)

    def reset(self, rng):
        ks = tf.split(rng, 3)
        qpos, qvel = self._sample_joints(ks[:, 1], ks[:, 2])
        qp = self.sys.default_qp(qpos, qvel)
        info = self.sys.info(qp)
        obs = np.concatenate(self._common_obs(qp, info, full_pos=False), axis=-1).astype(F)
        z = np.zeros(rng.shape[0], F)
        metrics = {k: z.copy() for k in ('reward_ctrl_cost', 'reward_contact_cost', 'reward_forward',
                                         'reward_survive')}
        return State(qp, obs, z.copy(), z.copy(), metrics, {})

    def step(self, state, action):
        action = np.asarray(action, F)
        qp, info = self.sys.step(state.qp, action)
        obs = np.concatenate(self._common_obs(qp, info, full_pos=False), axis=-1).astype(F)
        forward = (qp.pos[:, 0, 0] - state.qp.pos[:, 0, 0]) / self.sys.dt
        ctrl = F(.5) * np.sum(np.square(action), axis=-1)
        contact = F(0.5 * 1e-3) * np.sum(np.square(_clip1(info.contact_vel)).reshape(len(ctrl), -1), axis=-1)
        survive = np.ones_like(ctrl)
        reward = (forward - ctrl - contact + survive).astype(F)
        done = self._dead(qp)
        metrics = dict(state.metrics, reward_ctrl_cost=ctrl, reward_contact_cost=contact,
                       reward_forward=forward, reward_survive=survive)
        return state.replace(qp=qp, obs=obs, reward=reward, done=done, metrics=metrics)


class AntHeavenHellEnv(_AntBase):
    """ant_heavenhell.py:42-158."""

    def __init__(self, heaven_hell=((-5.25, 7.), (5.25, 7.)), priest_position=(0., 7.), visible_radius=2.,
                 dying_cost=-2., **kw):
        hhp = [tuple(heaven_hell[0]), tuple(heaven_hell[1]), tuple(priest_position)]
        super().__init__(cfgs.heavenhell_config(hhp=hhp, hallway_width=2.), **kw)
        self._hhp = np.concatenate([np.array(hhp, F), np.ones((3, 1), F)], axis=1)
        self.visible_radius, self.dying_cost = F(visible_radius), F(dying_cost)
        ix = self.sys.index
        self.target_idx, self.hell_idx, self.priest_idx = ix['Target'], ix['Hell'], ix['Priest']
        self.n_ant = self.priest_idx - self.torso_idx  # bodies 0..9: ant parts AND Ground (:70)
        self._init_lo, self._init_hi = np.array([-0.5, 0.5], F), np.array([0.5, 1.5], F)

    def sample_init_qp(self, rng):  # :87-103
        ks = tf.split(rng, 5)
        qpos, qvel = self._sample_joints(ks[:, 1], ks[:, 2])
        ant_pos = tf.uniform(ks[:, 3], 2, self._init_lo, self._init_hi)
        qp = self.sys.default_qp(qpos, qvel)
        qp.pos[:, :self.n_ant, :2] += ant_pos[:, None, :]
        order = tf.choice_no_replace(ks[:, 3], 2, 2)  # rng3 reused (:99), rng4 unused
        qp.pos[:, self.target_idx] = self._hhp[order[:, 0]]
        qp.pos[:, self.hell_idx] = self._hhp[order[:, 1]]
        return ks[:, 0], qp

    def reset(self, rng):  # :75-85
        rng, qp = self.sample_init_qp(rng)
        info = self.sys.info(qp)
        obs = self._get_obs(qp, info, np.zeros(rng.shape[0], bool))
        z = np.zeros(rng.shape[0], F)
        return State(qp, obs, z.copy(), z.copy(), {'heavens': z.copy(), 'hells': z.copy()}, {'rng': rng})

    def step(self, state, action):  # :106-123
        qp, info = self.sys.step(state.qp, np.asarray(action, F))
        dead = self._dead(qp)
        reward = np.where(dead > 0, self.dying_cost, F(0))
        txy = qp.pos[:, self.torso_idx, :2]
        rng_ok = [(_norm2(qp.pos[:, i, :2] - txy) <= self.visible_radius)
                  for i in (self.target_idx, self.hell_idx, self.priest_idx)]
        reward = np.where(rng_ok[0], F(1), reward)
        reward = np.where(rng_ok[1], F(-1), reward).astype(F)
        done = np.where(reward != 0, F(1), F(0)).astype(F)
        obs = self._get_obs(qp, info, rng_ok[2])
        metrics = dict(state.metrics, hits=done)
        return state.replace(qp=qp, obs=obs, reward=reward, done=done, metrics=metrics)

    def _get_obs(self, qp, info, priest_in_range):  # :125-158
        tgt_x = qp.pos[:, self.target_idx, 0]
        heaven_dir = np.where(priest_in_range, np.sign(tgt_x), F(0)).astype(F)
        return np.concatenate(self._common_obs(qp, info) + [heaven_dir[:, None]], axis=-1).astype(F)


class AntTagEnv(_AntBase):
    """ant_tag.py:28-181."""

    def __init__(self, tag_radius=1.5, visible_radius=3., target_step=0.5, min_spawn_distance=5.,
                 cage_xy=(4.5, 4.5), dying_cost=-1., **kw):
        super().__init__(cfgs.tag_config(cage_max_xy=cage_xy, offset=1.), **kw)
        self.tag_radius, self.visible_radius = F(tag_radius), F(visible_radius)
        self.target_step, self.min_spawn_distance = F(target_step), F(min_spawn_distance)
        self.cage_xy, self.dying_cost = np.array(cage_xy, F), F(dying_cost)
        self.target_idx = self.sys.index['Target']
        self.n_ant = self.target_idx - self.torso_idx  # 0..9 incl. Ground (:59)

    def _random_target(self, rng, ant_xy):  # :90-105 (batched while_loop)
        rng = rng.copy()
        xy = tf.uniform(rng, 2, -self.cage_xy, self.cage_xy)
        self.last_reject_iters = np.zeros(rng.shape[0], np.int32)
        while True:
            todo = _norm2(xy - ant_xy) <= self.min_spawn_distance
            if not todo.any():
                break
            k1 = tf.split(rng[todo], 2)[:, 1]
            xy[todo] = tf.uniform(k1, 2, -self.cage_xy, self.cage_xy)
            rng[todo] = k1
            self.last_reject_iters[todo] += 1
        return np.concatenate([xy, np.full((xy.shape[0], 1), F(0.5))], axis=1).astype(F)

    def reset(self, rng):  # :63-88
        ks = tf.split(rng, 5)
        qpos, qvel = self._sample_joints(ks[:, 1], ks[:, 2])
        ant_pos = tf.uniform(ks[:, 3], 2, -self.cage_xy, self.cage_xy)
        qp = self.sys.default_qp(qpos, qvel)
        qp.pos[:, :self.n_ant, :2] += ant_pos[:, None, :]
        qp.pos[:, self.target_idx] = self._random_target(ks[:, 4], ant_pos)
        info = self.sys.info(qp)
        obs = self._get_obs(qp, info)
        z = np.zeros(rng.shape[0], F)
        return State(qp, obs, z.copy(), z.copy(), {'hits': z.copy()}, {'rng': ks[:, 0]})

    def _step_target(self, rng, ant_xy, tgt_xy):  # :129-146
        ks = tf.split(rng, 2)
        choice = tf.randint(ks[:, 1], 0, 4)
        v = ant_xy - tgt_xy
        v = v / _norm2(v)[:, None]
        cands = np.stack([np.stack([v[:, 1], -v[:, 0]], -1), np.stack([-v[:, 1], v[:, 0]], -1), -v,
                          np.zeros_like(v)], axis=1)
        new = cands[np.arange(len(choice)), choice] * self.target_step + tgt_xy
        out = (np.abs(new) > self.cage_xy).any(axis=-1)
        new = np.where(out[:, None], tgt_xy, new).astype(F)
        self.last_choice = choice
        return ks[:, 0], np.concatenate([new, np.ones((len(new), 1), F)], axis=1)

    def step(self, state, action):  # :107-127
        qp, info = self.sys.step(state.qp, np.asarray(action, F))
        dead = self._dead(qp)
        reward = np.where(dead > 0, self.dying_cost, F(0))
        rng, tgt = self._step_target(state.info['rng'], qp.pos[:, self.torso_idx, :2],
                                     qp.pos[:, self.target_idx, :2])
        qp.pos[:, self.target_idx] = tgt
        obs = self._get_obs(qp, info)
        hit = np.where(_norm2(qp.pos[:, self.torso_idx, :2] - qp.pos[:, self.target_idx, :2]) <= self.tag_radius,
                       F(1), F(0)).astype(F)
        reward = np.where(hit > 0, F(1), reward).astype(F)
        done = np.logical_or(dead, hit)
        return state.replace(qp=qp, obs=obs, reward=reward, done=done, metrics=dict(state.metrics, hits=hit),
                             info=dict(state.info, rng=rng))

    def _get_obs(self, qp, info):  # :148-181
        txy = qp.pos[:, self.target_idx, :2]
        vis = _norm2(txy - qp.pos[:, self.torso_idx, :2]) <= self.visible_radius
        txy = np.where(vis[:, None], txy, F(0))
        return np.concatenate(self._common_obs(qp, info) + [txy], axis=-1).astype(F)


class AntGatherEnv(_AntBase):
    """ant_gather.py:42-213."""

    def __init__(self, n_apples=8, n_bombs=8, cage_xy=(6, 6), robot_object_spacing=2., catch_range=1.,
                 n_bins=10, sensor_range=6., sensor_span=np.pi, dying_cost=-10., **kw):
        super().__init__(cfgs.gather_config(cage_max_xy=cage_xy, offset=1., n_apples=n_apples, n_bombs=n_bombs),
                         **kw)
        self.n_apples, self.n_bombs, self.n_objects, self.n_bins = n_apples, n_bombs, n_apples + n_bombs, n_bins
        self.dying_cost, self.sensor_range, self.catch_range = F(dying_cost), F(sensor_range), F(catch_range)
        self.half_span = F(sensor_span / 2)
        self.bin_res = F((2 * (sensor_span / 2)) / n_bins)
        nb = self.sys.num_bodies
        self.obj = np.arange(nb - self.n_objects, nb)
        xs = np.arange(-cage_xy[0], cage_xy[0] + 1, dtype=F)
        ys = np.arange(-cage_xy[1], cage_xy[1] + 1, dtype=F)
        X, Y = np.meshgrid(xs, ys)  # 'xy' indexing: y-major, x fastest (:88)
        g = np.stack([X.ravel(), Y.ravel()], axis=1)
        g = g[np.sqrt(g[:, 0] ** 2 + g[:, 1] ** 2) > robot_object_spacing]
        self.grid = np.concatenate([g, np.zeros((len(g), 1), F)], axis=1).astype(F)
        self.waiting_area = (self.grid[-1] + self.sensor_range * 2).astype(F)

    def sample_init_qp(self, rng):  # :109-123
        ks = tf.split(rng, 4)
        qpos, qvel = self._sample_joints(ks[:, 1], ks[:, 2])
        qp = self.sys.default_qp(qpos, qvel)
        idx = tf.choice_no_replace(ks[:, 3], len(self.grid), self.n_objects)
        self.last_object_idx = idx
        opos = self.grid[idx].copy()
        opos[:, :self.n_apples, 2] = F(1.)
        qp.pos[:, self.obj] = opos
        return qp

    def _distances(self, qp):
        return _norm2(qp.pos[:, self.torso_idx, None, :2] - qp.pos[:, self.obj, :2])

    def reset(self, rng):  # :93-107
        qp = self.sample_init_qp(rng)
        info = self.sys.info(qp)
        obs = self._get_obs(qp, info, self._distances(qp))
        z = np.zeros(rng.shape[0], F)
        metrics = {'apples': z.copy(), 'bombs': z.copy(), 'objects': z.copy()}
        return State(qp, obs, z.copy(), z.copy(), metrics, {'rng': rng.copy()})

    def step(self, state, action):  # :125-150
        qp, info = self.sys.step(state.qp, np.asarray(action, F))
        dist = self._distances(qp)
        obs = self._get_obs(qp, info, dist)
        dead = self._dead(qp)
        reward = np.where(dead > 0, self.dying_cost, F(0))
        in_range = dist <= self.catch_range
        qp.pos[:, self.obj] = np.where(in_range[..., None], self.waiting_area, qp.pos[:, self.obj])
        ia, ib = in_range[:, :self.n_apples], in_range[:, self.n_apples:]
        reward = np.where(ia.any(-1) & (dead == 0), F(1), reward)
        reward = np.where(ib.any(-1) & (dead == 0), F(-1), reward).astype(F)
        all_gone = (qp.pos[:, self.obj] == self.waiting_area).all(axis=(1, 2))
        done = np.where(all_gone, F(1), dead).astype(F)
        metrics = dict(state.metrics, apples=ia.sum(-1).astype(np.int32), bombs=ib.sum(-1).astype(np.int32))
        return state.replace(qp=qp, obs=obs, reward=reward, done=done, metrics=metrics)

    def _get_readings(self, qp, dist):  # :152-181
        N = dist.shape[0]
        rot = qp.rot[:, self.torso_idx]
        e = np.broadcast_to(np.array([0, 1, 0, 0], F), rot.shape)
        o = bx.quat_mul(bx.quat_mul(rot, e), bx.quat_inv(rot))[:, 1:3]
        ori = np.arctan2(o[:, 1], o[:, 0])
        oxy = qp.pos[:, self.obj, :2]
        angles = (np.arctan2(oxy[..., 0], oxy[..., 1]) - ori[:, None]).astype(F)
        in_range = dist <= self.sensor_range
        bins = np.where((np.abs(angles) <= self.half_span) & in_range,
                        ((angles + self.half_span) / self.bin_res).astype(np.int32), np.int32(-1))
        bins[:, self.n_apples:] = np.where(bins[:, self.n_apples:] >= 0, bins[:, self.n_apples:] + self.n_apples, -1)
        inten = np.where(bins >= 0, F(1.) - (dist / self.sensor_range), F(0)).astype(F)
        readings = np.zeros((N, self.n_bins * 2), F)
        rows = np.arange(N)
        for k in range(self.n_objects):  # sequential scatter: last writer wins; -1 wraps to the last bin
            readings[rows, bins[:, k]] = inten[:, k]
        self.last_bins = bins
        return readings

    def _get_obs(self, qp, info, dist):  # :183-213
        return np.concatenate(self._common_obs(qp, info) + [self._get_readings(qp, dist)], axis=-1).astype(F)


ENVS = {'ant': AntEnv, 'ant_heavenhell': AntHeavenHellEnv, 'ant_gather': AntGatherEnv, 'ant_tag': AntTagEnv}


class EpisodeAutoReset:
    """create(name, episode_length, action_repeat, auto_reset, batch_size=N): ActionRepeat ->
    brax EpisodeWrapper(env, L, 1) -> VmapWrapper -> brax AutoResetWrapper (cached first state)."""

    def __init__(self, env, episode_length=1000, auto_reset=True):
        self.env, self.episode_length, self.auto_reset = env, episode_length, auto_reset

    def reset(self, rng):
        s = self.env.reset(rng)
        n = rng.shape[0]
        s.info['steps'] = np.zeros(n, F)
        s.info['truncation'] = np.zeros(n, F)
        if self.auto_reset:
            s.info['first_qp'] = s.qp.copy()
            s.info['first_obs'] = s.obs.copy()
        return s

    def step(self, state, action):
        info = dict(state.info)
        if self.auto_reset:
            info['steps'] = np.where(state.done > 0, F(0), info['steps']).astype(F)
            state = state.replace(done=np.zeros_like(state.done), info=info)
        s = self.env.step(state.replace(qp=state.qp.copy(), info=info), action)
        steps = (s.info['steps'] + F(1)).astype(F)
        over = steps >= self.episode_length
        donef = np.asarray(s.done, F)
        s.info['truncation'] = np.where(over, F(1) - donef, F(0)).astype(F)
        s.info['steps'] = steps
        done = np.where(over, np.ones_like(s.done), s.done)
        s = s.replace(done=done)
        if self.auto_reset:
            d = np.asarray(done, bool)
            f = s.info['first_qp']
            qp = bx.QP(*[np.where(d.reshape((-1,) + (1,) * (x.ndim - 1)), x, y) for x, y in
                         ((f.pos, s.qp.pos), (f.rot, s.qp.rot), (f.vel, s.qp.vel), (f.ang, s.qp.ang))])
            s = s.replace(qp=qp, obs=np.where(d[:, None], s.info['first_obs'], s.obs))
        return s


def create(env_name, episode_length=1000, action_repeat=1, auto_reset=True, **kwargs):
    return EpisodeAutoReset(ENVS[env_name](action_repeat=action_repeat, **kwargs), episode_length, auto_reset)


def gym_reset_keys(key, num_envs):
    """VmapGymWrapper._reset (wrappers.py:160-163): keys = split(key, N+1); (next gym key, env keys)."""
    ks = tf.split(np.asarray(key, np.uint32), num_envs + 1)
    return ks[0], ks[1:]


class RandomizedAutoResetNaive:
    """wrappers.py:30-52 (and :55-80, identical results) on top of Episode(env) without AutoReset:
    steps <- 0 where done; done <- 0; inner step; qp/obs <- reset(state.info['rng']) where done."""

    def __init__(self, env, episode_length=1000):
        self.inner = EpisodeAutoReset(env, episode_length, auto_reset=False)
        self.env = env

    def reset(self, rng):
        return self.inner.reset(rng)

    def step(self, state, action):
        info = dict(state.info)
        info['steps'] = np.where(np.asarray(state.done, F) > 0, F(0), info['steps']).astype(F)
        state = state.replace(done=np.zeros_like(np.asarray(state.done, F)), info=info)
        s = self.inner.step(state, action)
        self.step_margin = self.env.sys.margin   # test aid (brax_v1._note_margin): the step's, not the reset's below
        fresh = self.inner.reset(s.info['rng'])
        d = np.asarray(s.done, bool)
        qp = bx.QP(*[np.where(d.reshape((-1,) + (1,) * (x.ndim - 1)), x, y) for x, y in
                     ((fresh.qp.pos, s.qp.pos), (fresh.qp.rot, s.qp.rot), (fresh.qp.vel, s.qp.vel),
                      (fresh.qp.ang, s.qp.ang))])
        return s.replace(qp=qp, obs=np.where(d[:, None], fresh.obs, s.obs))
