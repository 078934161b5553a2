"""Parameter-container factories mirroring guided_diffusion/nn.py.

In this package these modules only hold parameters under the reference's names (so
checkpoints load unchanged); their math runs in the CUDA kernels:
GroupNorm32 (nn.py:17-19) -> adb_groupnorm, conv_nd (:22-32) -> adb_conv_igemm,
linear (:35-39) -> adb_linear, timestep_embedding (:103-121) -> adb_timestep_embedding.
"""
import torch.nn as nn

from . import ops


class GroupNorm32(nn.GroupNorm):
    pass


def conv_nd(dims, *args, **kwargs):
    if dims == 1:
        return nn.Conv1d(*args, **kwargs)
    elif dims == 2:
        return nn.Conv2d(*args, **kwargs)
    raise ValueError(f"unsupported dimensions: {dims}")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def zero_module(module):
    """nn.py:68-74."""
    for p in module.parameters():
        p.detach().zero_()
    return module


def normalization(channels):
    """nn.py:93-100."""
    return GroupNorm32(32, channels)


def timestep_embedding(timesteps, dim, max_period=10000):
    """nn.py:103-121 on the device (int64 timesteps)."""
    if max_period != 10000:
        raise NotImplementedError("max_period is fixed at 10000 as in every reference call site")
    return ops.timestep_embedding(timesteps.long().contiguous(), dim)
