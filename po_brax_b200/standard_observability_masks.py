"""Index sets into the 87-dim plain-Ant observation, the 'ant' rows of
/root/reference/po_brax/standard_observability_masks.py:7 (POSITION), :26 (VELOCITY), :62 (CFRC).
Intended use, as in the reference: `obs[..., POSITION['ant']]`. The other 13 brax envs of that file
are out of scope."""
import numpy as np

POSITION = {'ant': np.arange(0, 13)}    # torso z, torso quaternion, 8 joint angles
VELOCITY = {'ant': np.arange(13, 27)}   # torso vel, torso ang, 8 joint velocities
CFRC = {'ant': np.arange(27, 87)}       # clip(contact.vel) 10x3, clip(contact.ang) 10x3

# aliases used by the package root
POSITION_MASKS, VELOCITY_MASKS, EXTRA_INFO_MASKS = POSITION, VELOCITY, CFRC


def apply_mask(obs, mask):
    """obs[..., mask] for torch tensors or numpy arrays (mask: one of the index arrays above)."""
    try:
        import torch
        if isinstance(obs, torch.Tensor):
            return obs[..., torch.as_tensor(mask, device=obs.device, dtype=torch.long)]
    except ImportError:  # pragma: no cover
        pass
    return obs[..., mask]
