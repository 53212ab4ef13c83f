"""Gym-style adapters mirroring /root/reference/po_brax/envs/wrappers.py:126-262 on top of the fused env.

gym itself is not a dependency: the spaces are exposed as plain (low, high, shape) records. Observations,
rewards and dones are torch CUDA tensors (the reference returns device arrays as well).

Aliasing contract. The fused env updates a State's buffers in place, so the tensors of step t would be overwritten
by step t + 1. The reference returns fresh immutable arrays every step, and rollout code relies on that
(`rewards.append(r)`, a `done` read one step late). The gym adapters therefore return COPIES by default
(`copy=True`: obs, reward, done and the metrics rows are cloned on the device, one small launch each);
`copy=False` hands out the live buffers for loops that consume a step's outputs before the next step (they are
then valid until the next `step()` / `reset()`)."""
from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from .. import _lib, random as prandom
from .env import Env

Box = namedtuple('Box', ['low', 'high', 'shape', 'dtype'])


class _FunctionalWrapper:
    """brax.envs.env.Wrapper surface: forwards everything to the wrapped env."""

    def __init__(self, env: Env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, rng):
        return self.env.reset(rng)

    @property
    def unwrapped(self):
        return self.env.unwrapped


def _episode_head(state):
    """The head shared by the autoreset wrappers (wrappers.py:36-40): steps <- 0 where done; done <- 0."""
    b = state.buf
    b['steps'].mul_(1.0 - b['done'])
    b['done'].zero_()


class RandomizedAutoResetWrapperNaive(_FunctionalWrapper):
    """wrappers.py:30-52: where done, qp/obs <- a fresh `reset(state.info['rng'])` (only qp and obs are taken, so
    HeavenHell / Gather, whose info['rng'] never changes, re-reset to the identical state; Tag's key advances every
    step). Wrap an env built with auto_reset=False. Envs without info['rng'] (plain ant) raise KeyError, as in the
    reference."""

    def step(self, state, action):
        if self.env.auto_reset:
            raise RuntimeError('wrap an env created with auto_reset=False')
        if state.buf['rng'] is None:
            raise KeyError('rng')
        _episode_head(state)
        state = self.env.step(state, action)
        steps = state.buf['steps'].clone()  # the select leaves info['steps'] alone (it is zeroed at the next head)
        state = self.env.reset_where_done(state, state.buf['rng'])
        state.buf['steps'].copy_(steps)
        return state


class RandomizedAutoResetWrapperOnTerminal(RandomizedAutoResetWrapperNaive):
    """wrappers.py:55-80: same result as Naive; the reference merely skips the reset computation when no env is
    done (lax.cond). The fused reset kernel already only computes the done envs."""


class RandomizedAutoResetWrapperCached(_FunctionalWrapper):
    """wrappers.py:83-123: brax-style cached autoreset whose cached first state is refreshed every
    `n_steps_between_updates` calls from `reset(split(info['rng'])[1])`, info['rng'] <- split(...)[0].
    Wrap an env built with auto_reset=True (the cached select runs inside the step kernel)."""

    def __init__(self, env: Env, n_steps_between_updates: int = 200):
        super().__init__(env)
        self.n_steps_between_updates = n_steps_between_updates
        self.steps = 0

    def step(self, state, action):
        if not self.env.auto_reset:
            raise RuntimeError('wrap an env created with auto_reset=True')
        self.steps += 1
        if (self.steps % self.n_steps_between_updates) == 0:
            if state.buf['rng'] is None:
                raise KeyError('rng')
            rng, rng1 = self.env.split_pairs(state.buf['rng'])
            s = self.env.reset(rng1)
            state.buf['first_qp'].copy_(s.buf['qp'])
            if state.buf['first_aux'] is not None:
                state.buf['first_aux'].copy_(s.buf['aux'])
            state.buf['first_obs'].copy_(s.buf['obs'])
            state.buf['rng'].copy_(rng)
        return self.env.step(state, action)


class VmapGymWrapper:
    """wrappers.py:126-172: batched env behind the gym VectorEnv API; keys = split(key, num_envs + 1)."""

    def __init__(self, env: Env, batch_size: int, seed: int = 0, backend: Optional[str] = None, copy: bool = True):
        if batch_size != env.batch_size:
            raise ValueError('batch_size must equal env.batch_size')
        self._env = env
        self.copy = copy
        self.metadata = {'render.modes': ['human', 'rgb_array'],
                         'video.frames_per_second': 1 / env.sys.config.dt}   # wrappers.py:140
        self.num_envs = batch_size
        self.seed(seed)
        self.backend = backend
        self._state = None
        inf = float('inf')
        self.single_observation_space = Box(-inf, inf, (env.observation_size,), 'float32')
        self.observation_space = Box(-inf, inf, (batch_size, env.observation_size), 'float32')
        self.single_action_space = Box(-1.0, 1.0, (env.action_size,), 'float32')
        self.action_space = Box(-1.0, 1.0, (batch_size, env.action_size), 'float32')

    def seed(self, seed: int = 0):
        self._key = prandom.prng_key(seed)

    def _reset_keys(self):
        """keys = split(self._key, N + 1): keys[0] is the next gym key (host), keys[1:] the env keys (device)."""
        n = self.num_envs
        keys = self._env.split_keys(self._key, n + 1, first=1, count=n)
        nxt = prandom.split_at(self._key, n + 1, 0)
        return nxt, keys

    def _out(self, s):
        """(obs, reward, done, info) of a step: fresh tensors unless copy=False (module docstring)."""
        if not self.copy:
            return s.obs, s.reward, s.done, s.metrics
        return s.obs.clone(), s.reward.clone(), s.done.clone(), self._env._metrics_view(s.buf, s.buf['metrics'].clone())

    def reset(self):
        self._key, keys = self._reset_keys()
        self._state = self._env.reset(keys)
        return self._state.obs.clone() if self.copy else self._state.obs

    def step(self, action):
        self._state = self._env.step(self._state, action)
        return self._out(self._state)

    @property
    def unwrapped(self):
        return self._env


class AutoresetVmapGymWrapper(VmapGymWrapper):
    """wrappers.py:240-262: when any env is done, draw a fresh batch of keys from the stored gym key and
    reset exactly the done envs (qp/obs replaced, steps zeroed; reward/done/metrics/rng/truncation kept).

    `sync_free=True` (default): the gym key lives in a 4-word device buffer while stepping and the library makes the
    "any env done?" decision, the key draw and the key advance on the device (`pobrax_reset_where_done_chain`), so
    a step is two launches and no host round trip. Keys and results are identical to `sync_free=False`, which
    follows the reference literally (`done.any()` on the host, keys drawn by a split launch)."""

    def __init__(self, env: Env, batch_size: int, seed: int = 0, backend: Optional[str] = None, sync_free: bool = True,
                 cuda_graph: bool = False, copy: bool = True):
        self._chain = None   # int32[4] device tensor (gym key k0, k1, flag, spare) while the key lives on the device
        self._graph = None   # captured (step + reset_where_done_chain) of the current State's buffers
        super().__init__(env, batch_size, seed, backend, copy)
        self.sync_free = sync_free
        self.cuda_graph = cuda_graph

    @property
    def _key(self):
        if self._chain is not None:   # pull the gym key back from the device (one small copy; rare: reset / inspection)
            v = self._chain[:2].cpu().numpy().view(np.uint32)
            self._host_key, self._chain, self._graph = (int(v[0]), int(v[1])), None, None
        return self._host_key

    @_key.setter
    def _key(self, key):
        self._host_key, self._chain, self._graph = (int(key[0]), int(key[1])), None, None

    def reset(self):
        self._graph = None   # a reset allocates a new State: the captured launches point at the old buffers
        return super().reset()

    def _capture(self):
        """cuda_graph=True: the gym step (fused env step + device-side autoreset with the key chain, three launches)
        is captured once per State into a CUDA graph and replayed, so a step costs one graph launch on the host.
        Small batches are bound by the host path, not by the kernels. Results are identical:
        the same launches on the same buffers. The library's lazy per-kernel setup runs on a scratch state first
        (nothing but kernel launches may happen while capturing)."""
        env = self._env
        scratch = env.reset(env.split_keys((0, 0), self.num_envs + 1, first=1, count=self.num_envs))
        scratch = env.step(scratch, torch.zeros((self.num_envs, env.action_size), device=env.device))
        env.reset_where_done_chain(scratch, torch.zeros(4, dtype=torch.int32, device=env.device))
        self._act = torch.zeros((self.num_envs, env.action_size), dtype=torch.float32, device=env.device)
        torch.cuda.current_stream(env.device).synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            s = env.step(self._state, self._act)
            self._graph_state = env.reset_where_done_chain(s, self._chain)
        self._graph = g

    def step(self, action):
        if self.sync_free:
            if self._chain is None:
                words = np.array([self._host_key[0], self._host_key[1], 0, 0], dtype=np.uint32).view(np.int32)
                self._chain = torch.from_numpy(words).to(self._env.device)
            if self.cuda_graph:
                if self._graph is None:
                    self._capture()
                if not isinstance(action, torch.Tensor):
                    action = torch.as_tensor(np.asarray(action, np.float32))
                self._act.copy_(action.reshape(self._act.shape), non_blocking=True)
                self._graph.replay()
                self._state = s = self._graph_state
                return self._out(s)
            self._state = s = self._env.step(self._state, action)
            self._state = s = self._env.reset_where_done_chain(s, self._chain)
            return self._out(s)
        self._state = s = self._env.step(self._state, action)
        if bool(s.done.any()):
            self._key, keys = self._reset_keys()
            self._state = s = self._env.reset_where_done(s, keys)
        return self._out(s)


class AutoresetGymWrapper:
    """wrappers.py:232-237 over brax's GymWrapper: the UNBATCHED gym.Env view (create_gym_env(batch_size=None)).
    reset: key1, key2 = split(key); state = env.reset(key2); key <- key1 (brax GymWrapper.reset). step: when the
    episode ends the env is reset in full from the key chain (fresh info['rng'] too, unlike the batched adapter)
    and the RESET observation is returned together with the finished step's reward / done (`obs` is rebound by the
    reset in the reference). Wraps an un-vmapped env (create(batch_size=None): State fields carry no batch axis);
    `if done` is a host decision in the reference too. Outputs are copies (module docstring)."""

    def __init__(self, env: Env, seed: int = 0, backend: Optional[str] = None):
        if not env.unbatched:
            raise ValueError('AutoresetGymWrapper wraps an unbatched env (create(..., batch_size=None))')
        self._env = env
        self.metadata = {'render.modes': ['human', 'rgb_array'],
                         'video.frames_per_second': 1 / env.sys.config.dt}
        self.seed(seed)
        self.backend = backend
        self._state = None
        inf = float('inf')
        self.observation_space = Box(-inf, inf, (env.observation_size,), 'float32')
        self.action_space = Box(-1.0, 1.0, (env.action_size,), 'float32')

    def seed(self, seed: int = 0):
        self._key = prandom.prng_key(seed)

    def _reset(self):
        key1, key2 = prandom.split_at(self._key, 2, 0), prandom.split_at(self._key, 2, 1)
        self._state = self._env.reset(np.array(key2, dtype=np.uint32))
        self._key = key1
        return self._state.obs.clone()

    def reset(self):
        return self._reset()

    def step(self, action):
        if not isinstance(action, torch.Tensor):
            action = torch.as_tensor(np.asarray(action, np.float32))
        self._state = s = self._env.step(self._state, action.reshape(-1))
        obs, reward, done, info = s.obs.clone(), s.reward.clone(), s.done.clone(), {k: v.clone() for k, v in s.metrics.items()}
        if bool(done):
            obs = self._reset()
        return obs, reward, done, info

    @property
    def unwrapped(self):
        return self._env


class EvalGymWrapper:
    """wrappers.py:175-229: running episode statistics (returns, discounted returns, lengths). The reference
    appends finished episodes to Python lists and reports their nanmean; here the same means are kept as
    device-side sums and counts, so step() needs no host synchronisation."""

    def __init__(self, env, discount: float = 1.):
        self.env = env
        self._discount = discount
        self.num_envs = getattr(env, 'num_envs', 1)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kwargs):
        o = self.env.reset(**kwargs)
        z = torch.zeros(self.num_envs, dtype=torch.float32, device=o.device)
        self.episode_returns, self.discounted_episode_returns = z.clone(), z.clone()
        self.episode_lengths = torch.zeros(self.num_envs, dtype=torch.int64, device=o.device)
        self.current_discount = torch.ones_like(z)
        self._sums = torch.zeros(4, dtype=torch.float64, device=o.device)  # count, sum r, sum disc r, sum len
        return o

    def step(self, action):
        o, r, d, info = self.env.step(action)
        # one launch (pobrax_eval_update) instead of ~25 elementwise / reduction launches: same update, same sums
        rr = r.reshape(-1)
        dd = d.reshape(-1)
        if rr.dtype != torch.float32 or not rr.is_contiguous():
            rr = rr.float().contiguous()
        if dd.dtype != torch.float32 or not dd.is_contiguous():
            dd = dd.float().contiguous()
        idx = rr.device.index
        raw = getattr(torch._C, '_cuda_getCurrentRawStream', None)   # (torch.cuda.current_stream() costs 6 us a call)

        def launch():
            stream = raw(idx) if raw is not None else torch.cuda.current_stream(rr.device).cuda_stream
            _lib.check(_lib.load().pobrax_eval_update(
                rr.data_ptr(), dd.data_ptr(), self.episode_returns.data_ptr(), self.discounted_episode_returns.data_ptr(),
                self.episode_lengths.data_ptr(), self.current_discount.data_ptr(), self._sums.data_ptr(),
                float(self._discount), int(self.num_envs), stream), 'pobrax_eval_update')

        if torch.cuda.current_device() == idx:
            launch()
        else:   # the launch needs the tensors' device current
            with torch.cuda.device(idx):
                launch()
        return o, r, d, info

    def get_stats(self):
        c, sr, sd, sl = self._sums.tolist()
        nan = float('nan')
        return {'charts/mean_episodic_return': sr / c if c else nan,
                'charts/mean_discounted_episodic_return': sd / c if c else nan,
                'charts/mean_episodic_length': sl / c if c else nan}
