"""po_brax_b200: B200-native (sm_100a) fused stepper for po-brax's Ant POMDP envs behind the reference env API.

    from po_brax_b200 import envs
    env = envs.create('ant_heavenhell', batch_size=4096)
    state = env.reset(keys)            # keys: uint32 [N, 2] threefry keys
    state = env.step(state, action)    # action: float32 CUDA tensor [N, 8]
"""
from . import envs  # noqa: F401
from .standard_observability_masks import POSITION_MASKS, VELOCITY_MASKS, EXTRA_INFO_MASKS  # noqa: F401

__version__ = '0.1.0'
