"""Host-side mirror of the reference env API on top of the C-ABI (include/pobrax.h).

Mirrors, for the four Ant tasks, the interface of brax.envs.env.Env / State as used by
/root/reference/po_brax/envs/ant_heavenhell.py:42-158, ant_gather.py:42-213, ant_tag.py:28-181 and
brax.envs.ant.Ant: `reset(rng) -> State`, `step(state, action) -> State`,
`State(qp, obs, reward, done, metrics, info)`, `observation_size`, `action_size`, `unwrapped`.

All buffers are torch CUDA tensors owned by the State; the library only launches kernels on
torch's current stream. The env is always batched (the reference's VmapWrapper axis); the wrapper
stack ActionRepeat -> Episode -> Vmap -> AutoReset of `create()` is fused into the step kernel.
"""
import collections
import ctypes as C
import math
from typing import Dict, Optional

import numpy as np
import torch

_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)   # private but stable accessor (also what triton uses)

from .. import _lib

KINDS = {'ant': _lib.ANT, 'ant_heavenhell': _lib.ANT_HEAVENHELL, 'ant_gather': _lib.ANT_GATHER,
         'ant_tag': _lib.ANT_TAG}


_ANT_BODIES = ('$ Torso', 'Aux 1', '$ Body 4', 'Aux 2', '$ Body 7', 'Aux 3', '$ Body 10', 'Aux 4', '$ Body 13', 'Ground')


def _body_names(env_name, n_apples=8, n_bombs=8):
    """Body order of the reference systems (brax.envs.ant._SYSTEM_CONFIG + ant_heavenhell.py:17-32, ant_tag.py:16-24,
    ant_gather.py:25-38): qp rows and sys.body.index."""
    if env_name == 'ant':
        return _ANT_BODIES
    if env_name == 'ant_heavenhell':
        return _ANT_BODIES + ('Priest', 'Target', 'Hell', 'Arena')
    if env_name == 'ant_tag':
        return _ANT_BODIES + ('Target', 'Arena')
    return _ANT_BODIES + ('Arena',) + tuple(f'Target_{i + 1}' for i in range(n_apples)) + \
        tuple(f'Bomb_{i + 1}' for i in range(n_bombs))


class _Namespace:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class SysShim:
    """The part of `env.sys` (brax.System) that code outside the env reads: `sys.config.dt` / `.substeps`
    (/root/reference/po_brax/envs/wrappers.py:22-23,140), `sys.body.index[name]`, `sys.num_bodies`,
    `sys.num_joint_dof` (ant_heavenhell.py:64-69,90), `sys.default_angle()`. The physics itself (sys.step / info /
    default_qp) only exists fused inside the kernels: those attributes raise."""

    def __init__(self, env):
        p = env.params
        names = _body_names(env.env_name, p.n_apples, p.n_bombs)
        # ActionRepeatWrapper (wrappers.py:21-23) scales the config in place: dt *= k, substeps *= k
        self.config = _Namespace(dt=float(p.dt) * p.action_repeat, substeps=int(p.substeps) * p.action_repeat,
                                 bodies=[_Namespace(name=n) for n in names], gravity=(0.0, 0.0, float(p.gravity_z)),
                                 friction=float(p.friction), elasticity=float(p.elasticity),
                                 baumgarte_erp=float(p.baumgarte_erp), angular_damping=float(p.angular_damping),
                                 velocity_damping=float(p.velocity_damping))
        self.body = _Namespace(index={n: i for i, n in enumerate(names)})
        self.num_bodies = len(names)
        self.num_joints = 8
        self.num_joint_dof = 8
        self.num_forces_dof = 0
        self._env = env

    def default_angle(self):
        """System.default_angle(): midpoint of each joint's limit, radians, joint order hip 1, ankle 1, hip 2, ..."""
        p, out = self._env.params, []
        for leg in range(4):
            out.append((p.hip_limit[leg][0] + p.hip_limit[leg][1]) / 2 * math.pi / 180)
            out.append((p.ank_limit[leg][0] + p.ank_limit[leg][1]) / 2 * math.pi / 180)
        return torch.tensor(out, dtype=torch.float32, device=self._env.device)

    def __getattr__(self, name):
        if name in ('step', 'info', 'default_qp', 'joints', 'actuators', 'colliders'):
            raise AttributeError(f'sys.{name}: the rigid-body pipeline is fused into the CUDA step / reset kernels; '
                                 'use env.step / env.reset')
        raise AttributeError(name)


class EvalMetrics:
    """brax.envs.wrappers.EvalMetrics -- the record create(..., eval_metrics=True) puts into
    state.info['eval_metrics'] (__init__.py:69-70): current_episode_metrics {name: [N]}, completed_episodes_metrics
    {name: scalar}, completed_episodes, completed_episodes_steps. Plus `.acc` (all 8 device accumulators by name,
    dev_const.h) with dict-style access (`em['episodes']`, `em.items()`)."""
    __slots__ = ('current_episode_metrics', 'completed_episodes_metrics', 'completed_episodes',
                 'completed_episodes_steps', 'acc')

    def __init__(self, current_episode_metrics, completed_episodes_metrics, completed_episodes,
                 completed_episodes_steps, acc):
        self.current_episode_metrics, self.completed_episodes_metrics = current_episode_metrics, completed_episodes_metrics
        self.completed_episodes, self.completed_episodes_steps, self.acc = completed_episodes, completed_episodes_steps, acc

    def items(self):
        return self.acc.items()

    def __getitem__(self, k):
        return self.acc[k]


class QP:
    """brax.QP: pos[N,nb,3], rot[N,nb,4] (w,x,y,z), vel[N,nb,3], ang[N,nb,3] (torch tensors)."""
    __slots__ = ('pos', 'rot', 'vel', 'ang')

    def __init__(self, pos, rot, vel, ang):
        self.pos, self.rot, self.vel, self.ang = pos, rot, vel, ang

    def replace(self, **kw):
        d = dict(pos=self.pos, rot=self.rot, vel=self.vel, ang=self.ang)
        d.update(kw)
        return QP(**d)


_BUF_NAMES = ('qp', 'aux', 'obs', 'reward', 'done', 'steps', 'truncation', 'rng', 'metrics', 'first_qp',
              'first_aux', 'first_obs', 'ep_return', 'acc')


class State:
    """brax.envs.env.State over packed device buffers. `qp` / `info['first_qp']` are unpacked on access."""

    def __init__(self, env, buf: Dict[str, Optional[torch.Tensor]]):
        self._env = env
        self.buf = buf
        self._qp = None
        self._cstate = None
        self._stepped = False      # produced by env.step (not reset / state_from_qp)

    # ---- the reference's fields (an un-vmapped env, create(batch_size=None), has no leading axis)
    def _x(self, t):
        return t[0] if self._env.unbatched else t

    @property
    def qp(self) -> QP:
        if self._qp is None:
            q = self._env._unpack(self.buf['qp'], self.buf['aux'])
            self._qp = QP(q.pos[0], q.rot[0], q.vel[0], q.ang[0]) if self._env.unbatched else q
        return self._qp

    @property
    def obs(self):
        return self._x(self.buf['obs'])

    @property
    def reward(self):
        return self._x(self.buf['reward'])

    @property
    def done(self):
        """f32 0/1 like the reference -- except Tag after a step, where the reference's `done` is a bool
        (`jp.logical_or(dead, hit)`, ant_tag.py:127; f32 zeros at reset, :88)."""
        d = self.buf['done']
        if self._stepped and self._env.env_name == 'ant_tag':
            d = d != 0    # a fresh tensor on every access: the buffer is updated in place by later steps
        return self._x(d)

    @property
    def metrics(self) -> Dict[str, torch.Tensor]:
        m = self._env._metrics_view(self.buf)
        return {k: v[0] for k, v in m.items()} if self._env.unbatched else m

    @property
    def info(self) -> Dict[str, object]:
        b = self.buf
        info = {'steps': self._x(b['steps']), 'truncation': self._x(b['truncation'])}
        if b['rng'] is not None:
            info['rng'] = self._x(b['rng'])  # uint32 bit patterns stored as int32 [N, 2]
        if b['acc'] is not None:  # create(..., eval_metrics=True): brax EvalWrapper's record, kept on the device
            info['eval_metrics'] = self._env._eval_metrics(b)
        if b['first_qp'] is not None:
            info['first_obs'] = self._x(b['first_obs'])
            info['first_qp'] = _LazyQP(self._env, b['first_qp'], b['first_aux'])
        return info

    def replace(self, **kw):
        """State.replace for the fields the reference's wrappers overwrite (qp, obs, reward, done)."""
        buf = dict(self.buf)
        for k, v in kw.items():
            if k == 'qp':
                buf['qp'], buf['aux'] = self._env._pack(v, like_aux=self.buf['aux'])
            elif k in ('obs', 'reward', 'done'):
                v = v.to(torch.float32)
                buf[k] = (v.unsqueeze(0) if self._env.unbatched else v).contiguous()
            else:
                raise TypeError(f'State.replace: unsupported field {k!r}')
        out = State(self._env, buf)
        out._stepped = self._stepped
        return out

    def clone(self):
        out = State(self._env, {k: (None if v is None else v.clone()) for k, v in self.buf.items()})
        out._stepped = self._stepped
        return out

    def _c(self):
        if self._cstate is None:
            cs = _lib.PobraxState()
            for n in _BUF_NAMES:
                t = self.buf[n]
                setattr(cs, n, None if t is None else t.data_ptr())
            self._cstate = cs
        return self._cstate


class _LazyQP:
    def __init__(self, env, qp, aux):
        self._env, self._qp, self._aux, self._v = env, qp, aux, None

    def _get(self):
        if self._v is None:
            q = self._env._unpack(self._qp, self._aux)
            self._v = QP(q.pos[0], q.rot[0], q.vel[0], q.ang[0]) if self._env.unbatched else q
        return self._v

    pos = property(lambda s: s._get().pos)
    rot = property(lambda s: s._get().rot)
    vel = property(lambda s: s._get().vel)
    ang = property(lambda s: s._get().ang)


def _as_keys(rng, n, device):
    """[N,2] uint32 keys (numpy uint32 / torch int32|uint32|int64) -> contiguous int32 CUDA tensor."""
    if isinstance(rng, torch.Tensor):
        t = rng
        if t.dtype == torch.int64:
            t = (t & 0xFFFFFFFF).to(torch.int64)
            t = torch.where(t >= 2 ** 31, t - 2 ** 32, t).to(torch.int32)
        elif t.dtype in (torch.uint32,):
            t = t.view(torch.int32)
        elif t.dtype != torch.int32:
            raise TypeError(f'rng tensor must be int32/uint32/int64, got {t.dtype}')
    else:
        a = np.ascontiguousarray(np.asarray(rng).astype(np.uint32, copy=False))
        t = torch.from_numpy(a.view(np.int32))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.shape != (n, 2):
        raise ValueError(f'rng must have shape ({n}, 2) (one threefry key per env), got {tuple(t.shape)}')
    return t.to(device=device).contiguous()


class Env:
    """One batched po-brax Ant task bound to a CUDA device."""

    def __init__(self, env_name: str, batch_size: Optional[int] = None, episode_length: int = 1000,
                 action_repeat: int = 1, auto_reset: bool = True, device=None, track_metrics: bool = False,
                 **kwargs):
        if env_name not in KINDS:
            raise KeyError(env_name)
        if not torch.cuda.is_available():
            raise RuntimeError('po_brax_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.env_name = env_name
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('po_brax_b200 envs live on a CUDA device')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self._dev_index = int(self.device.index)
        # create(batch_size=None) adds no VmapWrapper (__init__.py:64): State fields carry no batch axis. The kernels
        # always run a batch; an unbatched env is a batch of one whose State strips the axis.
        self.unbatched = not batch_size    # `if batch_size:` in the reference: None and 0 both mean un-vmapped
        self.batch_size = 1 if self.unbatched else int(batch_size)
        if self.batch_size <= 0:
            raise ValueError('`batch_size` must be None or a positive integer')
        p = _lib.PobraxParams()
        _lib.check(self.lib.pobrax_default_params(KINDS[env_name], C.byref(p)), 'pobrax_default_params')
        p.num_envs = self.batch_size
        p.episode_length = int(episode_length) if episode_length else 0
        p.action_repeat = int(action_repeat)
        p.auto_reset = _lib.AUTORESET_CACHED if auto_reset else _lib.AUTORESET_OFF
        p.track_metrics = 1 if track_metrics else 0
        self._apply_kwargs(p, dict(kwargs))
        self.params = p
        L = _lib.PobraxLayout()
        _lib.check(self.lib.pobrax_layout(C.byref(p), C.byref(L)), 'pobrax_layout')
        self.layout = L
        h = C.c_void_p()
        _lib.check(self.lib.pobrax_create(C.byref(p), self.device.index, C.byref(h)), 'pobrax_create')
        self._h = h
        self.auto_reset = bool(auto_reset)
        self.track_metrics = bool(track_metrics)
        self.episode_length = p.episode_length
        self.sys = SysShim(self)
        ix = self.sys.body.index
        self.torso_idx = ix['$ Torso']
        for attr, name in (('target_idx', 'Target'), ('hell_idx', 'Hell'), ('priest_idx', 'Priest')):
            if name in ix:
                setattr(self, attr, ix[name])
        if env_name == 'ant_gather':   # ant_gather.py:77-86
            self.n_apples, self.n_bombs, self.n_bins = p.n_apples, p.n_bombs, p.n_bins
            self.n_objects = p.n_apples + p.n_bombs
            self.object_indices = list(range(self.num_bodies - self.n_objects, self.num_bodies))

    # ---- constructor kwargs of the reference envs
    def _apply_kwargs(self, p, kw):
        name = self.env_name

        def pop2(key, dst):
            if key in kw:
                v = kw.pop(key)
                dst[0], dst[1] = float(v[0]), float(v[1])

        if 'dying_cost' in kw and name != 'ant':
            p.dying_cost = float(kw.pop('dying_cost'))
        if name == 'ant_heavenhell':  # ant_heavenhell.py:51-56
            if 'heaven_hell' in kw:
                hh = kw.pop('heaven_hell')
                for i in range(2):
                    p.heaven_hell_xy[i][0], p.heaven_hell_xy[i][1] = float(hh[i][0]), float(hh[i][1])
            pop2('priest_position', p.priest_xy)
            if 'visible_radius' in kw:
                p.visible_radius = float(kw.pop('visible_radius'))
            if 'init_ant_pos' in kw:  # test hook: ant_heavenhell.py:73 self._init_ant_pos = [[lo_x, lo_y], [hi_x, hi_y]]
                ip = kw.pop('init_ant_pos')
                p.init_lo[0], p.init_lo[1], p.init_hi[0], p.init_hi[1] = (float(ip[0][0]), float(ip[0][1]),
                                                                          float(ip[1][0]), float(ip[1][1]))
            hw = 2.0  # ant_heavenhell.py:63 hallway_width
            xs = [p.heaven_hell_xy[0][0], p.heaven_hell_xy[1][0], p.priest_xy[0]]
            ys = [p.heaven_hell_xy[0][1], p.heaven_hell_xy[1][1], p.priest_xy[1]]
            _lib.check(self.lib.pobrax_draw_t_maze(C.byref(p), max(xs) + hw / 2, max(ys) + hw / 2, hw, 0.5),
                       'pobrax_draw_t_maze')
        elif name == 'ant_tag':  # ant_tag.py:38-45
            for k in ('tag_radius', 'visible_radius', 'target_step', 'min_spawn_distance'):
                if k in kw:
                    setattr(p, k, float(kw.pop(k)))
            pop2('cage_xy', p.cage_xy)
            p.init_lo[0], p.init_lo[1] = -p.cage_xy[0], -p.cage_xy[1]
            p.init_hi[0], p.init_hi[1] = p.cage_xy[0], p.cage_xy[1]
            _lib.check(self.lib.pobrax_draw_arena(C.byref(p), p.cage_xy[0] + 1.0, p.cage_xy[1] + 1.0, 0.5),
                       'pobrax_draw_arena')
        elif name == 'ant_gather':  # ant_gather.py:59-69
            for k in ('n_apples', 'n_bombs', 'n_bins'):
                if k in kw:
                    setattr(p, k, int(kw.pop(k)))
            for k in ('robot_object_spacing', 'catch_range', 'sensor_range', 'sensor_span'):
                if k in kw:
                    setattr(p, k, float(kw.pop(k)))
            pop2('cage_xy', p.gather_cage_xy)
            _lib.check(self.lib.pobrax_draw_arena(C.byref(p), p.gather_cage_xy[0] + 1.0, p.gather_cage_xy[1] + 1.0,
                                                  0.5), 'pobrax_draw_arena')
        if 'obs_mask' in kw:  # fused standard_observability_masks subset: a name ('position' | 'velocity' | 'cfrc',
            m = kw.pop('obs_mask')  # plain ant only) or an explicit (lo, hi) column range
            if isinstance(m, str):
                if name != 'ant':
                    raise ValueError('named observability masks exist for the plain ant only; pass (lo, hi)')
                m = {'position': (0, 13), 'velocity': (13, 27), 'cfrc': (27, 87)}[m]
            if m is not None:
                p.obs_col_lo, p.obs_col_hi = int(m[0]), int(m[1])
        if kw.pop('walls', True) is False:  # test hook: drop the Arena colliders
            p.num_walls = 0
        kw.pop('legacy_spring', None)
        if 'sys_dt' in kw:        # test hooks: brax config dt / substeps (sys.config.dt, sys.config.substeps)
            p.dt = float(kw.pop('sys_dt'))
        if 'sys_substeps' in kw:
            p.substeps = int(kw.pop('sys_substeps'))
        if kw:
            raise TypeError(f'{name}: unexpected constructor arguments {sorted(kw)}')

    def __del__(self):
        h = getattr(self, '_h', None)
        if h is not None and h.value:
            try:
                self.lib.pobrax_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---- brax.envs.env.Env surface
    @property
    def observation_size(self) -> int:
        return int(self.layout.obs_dim)

    @property
    def action_size(self) -> int:
        return int(self.layout.action_dim)

    @property
    def num_bodies(self) -> int:
        return int(self.layout.num_bodies)

    @property
    def unwrapped(self):
        return self

    def _stream(self):
        # the raw cudaStream_t of torch's current stream on this env's device; torch.cuda.current_stream() builds a
        # Stream object per call (6 us: a third of a small-batch step's host time), the C accessor does not
        if _raw_stream is not None:
            return C.c_void_p(_raw_stream(self._dev_index))
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _alloc(self) -> Dict[str, Optional[torch.Tensor]]:
        n, L, dev = self.batch_size, self.layout, self.device
        f = dict(dtype=torch.float32, device=dev)
        buf = {
            'qp': torch.zeros((L.qp_planes, n, 4), **f),
            'aux': torch.zeros((L.aux_dim, n), **f) if L.aux_dim else None,
            'obs': torch.empty((n, L.obs_dim), **f),
            'reward': torch.empty(n, **f), 'done': torch.empty(n, **f), 'steps': torch.empty(n, **f),
            'truncation': torch.empty(n, **f),
            'rng': torch.empty((n, 2), dtype=torch.int32, device=dev) if self.env_name != 'ant' else None,
            'metrics': torch.empty((L.metrics_dim, n), **f),
            'first_qp': None, 'first_aux': None, 'first_obs': None, 'ep_return': None, 'acc': None,
        }
        if self.auto_reset:
            buf['first_qp'] = torch.zeros((L.qp_planes, n, 4), **f)
            buf['first_aux'] = torch.zeros((L.aux_dim, n), **f) if L.aux_dim else None
            buf['first_obs'] = torch.empty((n, L.obs_dim), **f)
        if self.track_metrics:
            buf['ep_return'] = torch.zeros(n, **f)
            buf['acc'] = torch.zeros(_lib.NUM_ACC, dtype=torch.float64, device=dev)
        return buf

    def reset(self, rng) -> State:
        """env.reset(rng): rng = one threefry key per env, uint32 [N, 2]."""
        keys = _as_keys(rng, self.batch_size, self.device)
        st = State(self, self._alloc())
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pobrax_reset(self._h, keys.data_ptr(), C.byref(st._c()), self._stream()),
                       'pobrax_reset')
        st._keys = keys  # keep alive until the launch has been enqueued on this stream
        return st

    def step(self, state: State, action: torch.Tensor) -> State:
        """env.step(state, action). The state's buffers are updated in place (the reference also mutates
        state.metrics / state.info in place); the returned State shares them."""
        if not isinstance(action, torch.Tensor):
            action = torch.as_tensor(np.asarray(action, np.float32))
        if action.device != self.device or action.dtype != torch.float32 or not action.is_contiguous():
            action = action.to(device=self.device, dtype=torch.float32).contiguous()
        if self.unbatched and action.dim() == 1:
            action = action.unsqueeze(0)
        if action.shape != (self.batch_size, self.action_size):
            raise ValueError(f'action must have shape ({self.batch_size}, {self.action_size}), got {tuple(action.shape)}')
        # no torch.cuda.device() context here: the library selects the handle's device itself (small batches are
        # bound by this host path, not by the kernel)
        _lib.check(self.lib.pobrax_step(self._h, C.byref(state._c()), action.data_ptr(), self._stream()), 'pobrax_step')
        out = State(self, state.buf)
        out._cstate = state._cstate
        out._action = action
        out._stepped = True
        return out

    def reset_where_done(self, state: State, rng) -> State:
        """Gym-level autoreset (wrappers.py:245-262): where done, qp/obs <- reset(rng[i]) and steps <- 0."""
        keys = _as_keys(rng, self.batch_size, self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pobrax_reset_where_done(self._h, keys.data_ptr(), C.byref(state._c()), self._stream()),
                       'pobrax_reset_where_done')
        out = State(self, state.buf)
        out._cstate = state._cstate
        out._keys = keys
        out._stepped = state._stepped
        return out

    def reset_where_done_chain(self, state: State, chain: torch.Tensor) -> State:
        """reset_where_done with the gym key chain on the device: `chain` = int32[4] tensor (gym key k0, k1, flag = 0,
        spare). Where some env is done: keys = split(gym key, N + 1), done envs <- reset(keys[i + 1]), gym key <-
        keys[0] -- what AutoresetVmapGymWrapper.step (wrappers.py:245-262) does, without its host round trip."""
        if chain.dtype != torch.int32 or chain.numel() != 4 or chain.device != self.device or not chain.is_contiguous():
            raise ValueError('chain must be a contiguous int32[4] tensor on the env\'s device')
        _lib.check(self.lib.pobrax_reset_where_done_chain(self._h, chain.data_ptr(), C.byref(state._c()),
                                                          self._stream()), 'pobrax_reset_where_done_chain')
        out = State(self, state.buf)
        out._cstate = state._cstate
        out._stepped = state._stepped
        return out

    def split_keys(self, key, n=None, first=0, count=None) -> torch.Tensor:
        """jax.random.split(key, n)[first:first+count] on the device (int32 bit patterns)."""
        n = self.batch_size + 1 if n is None else n
        count = n - first if count is None else count
        k = (C.c_uint32 * 2)(int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF)
        out = torch.empty((count, 2), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pobrax_split_keys(k, n, first, count, out.data_ptr(), self._stream()),
                       'pobrax_split_keys')
        return out

    def split_pairs(self, keys: torch.Tensor):
        """vmapped jax.random.split(keys, 2) -> (keys[:, 0], keys[:, 1]) for device keys [N, 2]."""
        keys = keys.contiguous()
        a, b = torch.empty_like(keys), torch.empty_like(keys)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pobrax_split_pairs(keys.data_ptr(), keys.shape[0], a.data_ptr(), b.data_ptr(),
                                                   self._stream()), 'pobrax_split_pairs')
        return a, b

    # ---- helpers for State
    def _unpack(self, qp, aux) -> QP:
        n, nb = self.batch_size, self.num_bodies
        f = dict(dtype=torch.float32, device=self.device)
        pos, rot = torch.empty((n, nb, 3), **f), torch.empty((n, nb, 4), **f)
        vel, ang = torch.empty((n, nb, 3), **f), torch.empty((n, nb, 3), **f)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pobrax_unpack_qp(self._h, qp.data_ptr(), None if aux is None else aux.data_ptr(),
                                                 pos.data_ptr(), rot.data_ptr(), vel.data_ptr(), ang.data_ptr(),
                                                 self._stream()), 'pobrax_unpack_qp')
        return QP(pos, rot, vel, ang)

    def _pack(self, qp: QP, like_aux=None):
        n, nb, L = self.batch_size, self.num_bodies, self.layout
        f = dict(dtype=torch.float32, device=self.device)
        arrs = []
        for name, w in (('pos', 3), ('rot', 4), ('vel', 3), ('ang', 3)):
            t = getattr(qp, name)
            if not isinstance(t, torch.Tensor):
                t = torch.as_tensor(np.asarray(t, np.float32))
            t = t.to(**f)
            if self.unbatched and t.dim() == 2:
                t = t.unsqueeze(0)
            t = t.contiguous()
            if t.shape != (n, nb, w):
                raise ValueError(f'qp.{name} must have shape ({n}, {nb}, {w}), got {tuple(t.shape)}')
            arrs.append(t)
        if self.env_name == 'ant_heavenhell':
            # The packed state keeps one flag for the goal side: a QP whose Target / Hell / Priest rows are anything
            # but the configured positions cannot be represented -- refuse it instead of snapping silently.
            p = self.params
            hh = torch.tensor([[p.heaven_hell_xy[i][0], p.heaven_hell_xy[i][1]] for i in range(2)], **f)
            pr = torch.tensor([p.priest_xy[0], p.priest_xy[1]], **f)
            tgt, hell = arrs[0][:, self.target_idx, :2], arrs[0][:, self.hell_idx, :2]
            ok = (((tgt == hh[0]).all(-1) & (hell == hh[1]).all(-1)) | ((tgt == hh[1]).all(-1) & (hell == hh[0]).all(-1))) \
                & (arrs[0][:, self.priest_idx, :2] == pr).all(-1)
            if not bool(ok.all()):
                raise ValueError('ant_heavenhell: qp.pos rows Target / Hell must be the two configured heaven_hell '
                                 'positions (in either order) and Priest the configured priest_position')
        out = torch.zeros((L.qp_planes, n, 4), **f)
        aux = (like_aux.clone() if like_aux is not None else torch.zeros((L.aux_dim, n), **f)) if L.aux_dim else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pobrax_pack_qp(self._h, arrs[0].data_ptr(), arrs[1].data_ptr(), arrs[2].data_ptr(),
                                               arrs[3].data_ptr(), out.data_ptr(),
                                               None if aux is None else aux.data_ptr(), self._stream()),
                       'pobrax_pack_qp')
        return out, aux

    def state_from_qp(self, qp: QP, rng=None, steps=None) -> State:
        """Builds a State around a given brax QP (tests / teacher forcing); obs/reward/done are zeroed."""
        buf = self._alloc()
        buf['qp'], buf['aux'] = self._pack(qp)
        for k in ('obs', 'reward', 'done', 'steps', 'truncation', 'metrics'):
            buf[k].zero_()
        if buf['first_qp'] is not None:
            buf['first_qp'].copy_(buf['qp'])
            if buf['aux'] is not None:
                buf['first_aux'].copy_(buf['aux'])
            buf['first_obs'].zero_()
        if buf['rng'] is not None:
            buf['rng'].copy_(_as_keys(rng, self.batch_size, self.device) if rng is not None else
                             torch.zeros_like(buf['rng']))
        if steps is not None:
            buf['steps'].copy_(torch.as_tensor(np.asarray(steps, np.float32)).to(self.device))
        return State(self, buf)

    _METRIC_ROWS = {'ant': ('reward_ctrl_cost', 'reward_contact_cost', 'reward_forward', 'reward_survive'),
                    'ant_heavenhell': ('hits',), 'ant_tag': ('hits',), 'ant_gather': ('apples', 'bombs')}
    _METRIC_ZERO = {'ant': (), 'ant_heavenhell': ('heavens', 'hells'), 'ant_tag': (), 'ant_gather': ('objects',)}

    def _metrics_view(self, buf, rows=None):
        rows = buf['metrics'] if rows is None else rows
        m = {name: rows[i] for i, name in enumerate(self._METRIC_ROWS[self.env_name])}
        for name in self._METRIC_ZERO[self.env_name]:  # keys the reference creates and never updates
            m[name] = torch.zeros_like(buf['reward'])
        return m

    # which device accumulator holds the sum over completed episodes of each per-step metric (dev_const.h `acc`)
    _ACC_OF_METRIC = {'ant_heavenhell': {'hits': 4}, 'ant_tag': {'hits': 4}, 'ant_gather': {'apples': 4, 'bombs': 5},
                      'ant': {}}

    def _eval_metrics(self, buf) -> EvalMetrics:
        """brax EvalWrapper's record from the device-side accumulators (no per-step host work): per-env running
        return of the current episode, sums over completed episodes, their count and total length. Unlike brax
        0.0.12 -- whose current_episode_metrics restart from the terminal step's metrics instead of zero -- an
        episode's return here is exactly the sum of its own rewards."""
        acc = buf['acc']
        done_sums = {'reward': acc[1]}
        for name, i in self._ACC_OF_METRIC[self.env_name].items():
            done_sums[name] = acc[i]
        return EvalMetrics(current_episode_metrics={'reward': buf['ep_return'][0] if self.unbatched else buf['ep_return']},
                           completed_episodes_metrics=done_sums, completed_episodes=acc[0],
                           completed_episodes_steps=acc[2], acc=dict(zip(Env.ACC_NAMES, acc.unbind(0))))

    ACC_NAMES = ('episodes', 'sum_return', 'sum_length', 'truncations', 'hits_or_apples', 'heavens_or_bombs',
                 'hells', 'dead_steps')

