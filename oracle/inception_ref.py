"""Inception-V3 pool_3 reference graph in plain torch (test oracle; see oracle/__init__.py).

torchvision's `inception_v3` executed by PyTorch in fp32, optionally turned into the FID Inception the references use:
the 2015 TF graph of evaluations/evaluator_v1.py:252-280, 665-679, which pytorch-fid's InceptionV3
("Stable Diffusion"/scripts/search_ea.py:45, 95-127, 171-182) reproduces on top of torchvision by changing three pooling
details - the 3x3 average pools of the Mixed_5x / Mixed_6x / Mixed_7b blocks exclude the zero padding from the divisor,
and Mixed_7c's pool branch is a MAX pool - and by resizing with F.interpolate(bilinear, align_corners=False) and
normalising to [-1, 1]. This is what `autodiffusion_b200.inception.InceptionPool3` (hand-written gather / pooling
kernels + tcgen05 GEMMs, BatchNorm folded) is checked against on identical seeded weights: tests/test_inception_gpu.py.
No Inception weights exist offline, so parity is pinned at the activation level on random (He) weights; the same
state_dict layout (torchvision's) loads pytorch-fid's `pt_inception-2015-12-05` when it is available.
"""
from __future__ import annotations

from typing import Optional

import torch as th
import torch.nn as nn
import torch.nn.functional as F


def _fid_forward_a(self, x):  # InceptionA with count_include_pad=False
    b1 = self.branch1x1(x)
    b5 = self.branch5x5_2(self.branch5x5_1(x))
    b3 = self.branch3x3dbl_3(self.branch3x3dbl_2(self.branch3x3dbl_1(x)))
    bp = self.branch_pool(F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False))
    return [b1, b5, b3, bp]


def _fid_forward_c(self, x):  # InceptionC with count_include_pad=False
    b1 = self.branch1x1(x)
    b7 = self.branch7x7_3(self.branch7x7_2(self.branch7x7_1(x)))
    bd = self.branch7x7dbl_5(self.branch7x7dbl_4(self.branch7x7dbl_3(self.branch7x7dbl_2(self.branch7x7dbl_1(x)))))
    bp = self.branch_pool(F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False))
    return [b1, b7, bd, bp]


def _fid_forward_e(pool):
    def fwd(self, x):  # InceptionE; pool = padding-excluding average (Mixed_7b) or max (Mixed_7c)
        b1 = self.branch1x1(x)
        b3 = self.branch3x3_1(x)
        b3 = th.cat([self.branch3x3_2a(b3), self.branch3x3_2b(b3)], 1)
        bd = self.branch3x3dbl_2(self.branch3x3dbl_1(x))
        bd = th.cat([self.branch3x3dbl_3a(bd), self.branch3x3dbl_3b(bd)], 1)
        bp = self.branch_pool(pool(x))
        return [b1, b3, bd, bp]

    return fwd


def _apply_fid_variant(net):
    import types

    for name in ("Mixed_5b", "Mixed_5c", "Mixed_5d"):
        m = getattr(net, name)
        m._forward = types.MethodType(_fid_forward_a, m)
    for name in ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e"):
        m = getattr(net, name)
        m._forward = types.MethodType(_fid_forward_c, m)
    avg = lambda x: F.avg_pool2d(x, kernel_size=3, stride=1, padding=1, count_include_pad=False)
    mx = lambda x: F.max_pool2d(x, kernel_size=3, stride=1, padding=1)
    net.Mixed_7b._forward = types.MethodType(_fid_forward_e(avg), net.Mixed_7b)
    net.Mixed_7c._forward = types.MethodType(_fid_forward_e(mx), net.Mixed_7c)


class InceptionPool3Ref(nn.Module):
    def __init__(self, weights: Optional[str] = None, seed: int = 0, half: Optional[bool] = None, fid_variant: bool = True):
        super().__init__()
        import torchvision

        with th.random.fork_rng(devices=[]):
            th.manual_seed(seed)
            net = torchvision.models.inception_v3(weights=None, aux_logits=False, transform_input=False, init_weights=False)
            if weights is None:
                # fan-in scaled (He) draws keep the 94 conv layers' activations O(1); torchvision's own init (std 0.1
                # everywhere) makes an untrained network's features ~1e12
                for m in net.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight, nonlinearity="relu")
        net.fc = nn.Identity()
        if fid_variant:
            _apply_fid_variant(net)
        if weights is not None:
            sd = th.load(weights, map_location="cpu")
            missing, unexpected = net.load_state_dict(sd, strict=False)
            if any(not k.startswith("fc.") and not k.startswith("AuxLogits.") for k in list(missing) + list(unexpected)):
                raise ValueError(f"Inception weights do not fit torchvision's inception_v3: missing {missing}, unexpected {unexpected}")
        self.net = net.eval()
        # fp16 autocast only with trained weights: a randomly initialised Inception's activations overflow fp16
        self.half = (weights is not None) if half is None else half
        self.dim = 2048

    @th.no_grad()
    def forward(self, u8: th.Tensor) -> th.Tensor:
        """u8: uint8 [B, H, W, 3] (the sampler's packed images) -> fp32 [B, 2048]."""
        if u8.dtype != th.uint8 or u8.dim() != 4 or u8.shape[3] != 3:
            raise ValueError("InceptionPool3Ref expects uint8 NHWC RGB images")
        x = u8.permute(0, 3, 1, 2).float()
        x = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)
        x = x / 127.5 - 1.0  # pytorch-fid's 2 * (x / 255) - 1
        if self.half and x.is_cuda:
            with th.autocast("cuda", dtype=th.float16):
                f = self.net(x)
        else:
            f = self.net(x)
        f = f.float().reshape(u8.shape[0], -1)
        if not bool(th.isfinite(f).all()):
            raise FloatingPointError("InceptionPool3Ref produced non-finite features (fp16 overflow? construct with half=False)")
        return f
