"""Deterministic, architecture-independent weight recipe shared by the golden generator and the tests.

The reference's random init cannot be used as-is: `zero_module` zeroes every ResBlock's last
conv, every attention `proj_out` and the final `out` conv (guided_diffusion/nn.py:68-74,
dynamic_unet.py:219-221,311,653), so an untouched random-init model outputs exactly 0 and
parity would be vacuous. Instead every parameter is drawn from a CPU
generator seeded by (seed, crc32(name)) — independent of construction order — with fan-in
scaling so activations stay O(1) through 58 blocks.
"""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch


def make_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        is_norm = (".in_layers.0." in name or ".out_layers.0." in name or ".norm." in name or name.startswith("out.0."))
        if name.endswith(".bias"):
            t = (0.05 if is_norm else 0.02) * torch.randn(shape, generator=g)
        elif is_norm:
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name == "label_emb.weight":
            t = 0.3 * torch.randn(shape, generator=g)
        elif name.endswith("positional_embedding"):
            t = torch.randn(shape, generator=g) / shape[0] ** 0.5
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) / fan_in ** 0.5
        sd[name] = t.float()
    return sd
