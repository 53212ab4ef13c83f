"""Factory mirroring /root/reference/po_brax/envs/__init__.py:28-121 for the Ant-family envs."""
import functools
from typing import Callable, Optional

from .env import Env, QP, State  # noqa: F401

# __init__.py:29-33 (the 13 stock brax envs of :34-46 are out of scope)
_envs = {'ant': 'ant', 'ant_tag': 'ant_tag', 'ant_heavenhell': 'ant_heavenhell', 'ant_gather': 'ant_gather'}


def create(env_name: str, episode_length: int = 1000, action_repeat: int = 1, auto_reset: bool = True,
           batch_size: Optional[int] = None, eval_metrics: bool = False, **kwargs) -> Env:
    """__init__.py:50-72. The wrapper stack ActionRepeat -> Episode(episode_length, 1) -> Vmap -> AutoReset
    is fused into the step kernel; eval_metrics turns on the device-side episode accumulators."""
    if env_name not in _envs:
        raise KeyError(f'{env_name!r}: po_brax_b200 provides {sorted(_envs)}')
    return Env(_envs[env_name], batch_size=batch_size, episode_length=episode_length or 0,
               action_repeat=1 if action_repeat is None else action_repeat, auto_reset=auto_reset,
               track_metrics=eval_metrics, **kwargs)


def create_fn(env_name: str, **kwargs) -> Callable[..., Env]:
    """__init__.py:75-77."""
    return functools.partial(create, env_name, **kwargs)


def create_gym_env(env_name: str, batch_size: Optional[int] = None, seed: int = 0, backend: Optional[str] = None,
                   **kwargs):
    """__init__.py:98-121: autoreset and statistics move to the gym layer."""
    from .wrappers import AutoresetGymWrapper, AutoresetVmapGymWrapper, EvalGymWrapper
    kwargs['auto_reset'] = False
    eval_metrics = kwargs.pop('eval_metrics', False)
    discount = kwargs.pop('discount', 1.)
    cuda_graph = kwargs.pop('cuda_graph', False)   # extension: replay the gym step as one CUDA graph (wrappers.py)
    copy = kwargs.pop('copy', True)                # False: hand out the live device buffers (wrappers.py docstring)
    if batch_size is not None and batch_size <= 0:
        raise ValueError('`batch_size` should either be None or a positive integer.')
    environment = create(env_name=env_name, batch_size=batch_size, **kwargs)
    if batch_size is None:
        e = AutoresetGymWrapper(environment, seed=seed, backend=backend)
    else:
        e = AutoresetVmapGymWrapper(environment, batch_size, seed=seed, backend=backend, cuda_graph=cuda_graph,
                                    copy=copy)
    if eval_metrics:
        e = EvalGymWrapper(e, discount=discount)
    return e
