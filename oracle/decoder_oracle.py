"""TEST INFRASTRUCTURE ONLY -- CPU restatement of NeuRADModel.rgb_decoder.

The camera half of NeuRADModel.decode_features (models/neurad.py:359-366): the rendered feature image [B, H, W, 48]
goes through  Conv2d(48->32, 1x1) + ReLU -> 2 x BasicBlock(32, 7x7, pad 3, BatchNorm) -> ConvTranspose2d(32->32,
kernel = stride = 3) -> 2 x BasicBlock -> Conv2d(32->3, 1x1) -> Sigmoid  (models/neurad.py:201-216,
model_components/cnns.py:19-46), in eval mode (BatchNorm uses its running statistics).  Plain torch fp32 on the CPU;
pinned bit for bit against the reference's own nn.Sequential by oracle/make_golden_decoder.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F
from torch import Tensor

BN_EPS = 1e-5  # torch.nn.BatchNorm2d default, not overridden by cnns.py


def basic_block(p: Dict[str, Tensor], prefix: str, x: Tensor) -> Tensor:
    """BasicBlock.forward (model_components/cnns.py:19-46): relu(x + bn(conv(relu(bn(conv(x))))))."""
    h = x
    for conv, bn, relu in ((0, 1, True), (3, 4, False)):
        h = F.conv2d(h, p[f"{prefix}.main_branch.{conv}.weight"], p[f"{prefix}.main_branch.{conv}.bias"], padding=3)
        h = F.batch_norm(h, p[f"{prefix}.main_branch.{bn}.running_mean"], p[f"{prefix}.main_branch.{bn}.running_var"],
                         p[f"{prefix}.main_branch.{bn}.weight"], p[f"{prefix}.main_branch.{bn}.bias"], False, 0.0, BN_EPS)
        if relu:
            h = torch.relu(h)
    return torch.relu(x + h)


def rgb_decoder(p: Dict[str, Tensor], features: Tensor, upsample: int = 3, prefix: str = "rgb_decoder") -> Tensor:
    """features [B, H, W, C_in] (row-major rays) -> rgb [B, H*upsample, W*upsample, 3], as decode_features does
    (neurad.py:362-365: permute to NCHW, run the Sequential, permute back)."""
    x = features.permute(0, 3, 1, 2)
    x = torch.relu(F.conv2d(x, p[f"{prefix}.0.weight"], p[f"{prefix}.0.bias"]))
    x = basic_block(p, f"{prefix}.2", x)
    x = basic_block(p, f"{prefix}.3", x)
    x = F.conv_transpose2d(x, p[f"{prefix}.4.weight"], p[f"{prefix}.4.bias"], stride=upsample)
    x = basic_block(p, f"{prefix}.5", x)
    x = basic_block(p, f"{prefix}.6", x)
    x = torch.sigmoid(F.conv2d(x, p[f"{prefix}.7.weight"], p[f"{prefix}.7.bias"]))
    return x.permute(0, 2, 3, 1)


def random_decoder_params(seed: int, in_dim: int = 48, hidden: int = 32, upsample: int = 3, prefix: str = "rgb_decoder") -> Dict[str, Tensor]:
    """Random-init parameters with the reference's state_dict keys/shapes (torch default inits; BatchNorm statistics and
    affine parameters randomised so that folding them is actually exercised)."""
    g = torch.Generator().manual_seed(seed)

    def conv(co, ci, k):
        bound = 1.0 / (ci * k * k) ** 0.5
        return (torch.rand(co, ci, k, k, generator=g) * 2 - 1) * bound, (torch.rand(co, generator=g) * 2 - 1) * bound

    p: Dict[str, Tensor] = {}
    p[f"{prefix}.0.weight"], p[f"{prefix}.0.bias"] = conv(hidden, in_dim, 1)
    for blk in (2, 3, 5, 6):
        for c, b in ((0, 1), (3, 4)):
            w, bias = conv(hidden, hidden, 7)
            # gain ~sqrt(3): keeps the activations O(1) through 8 conv layers instead of decaying to the biases
            p[f"{prefix}.{blk}.main_branch.{c}.weight"], p[f"{prefix}.{blk}.main_branch.{c}.bias"] = w * 1.7, bias
            p[f"{prefix}.{blk}.main_branch.{b}.weight"] = torch.rand(hidden, generator=g) * 0.8 + 0.6
            p[f"{prefix}.{blk}.main_branch.{b}.bias"] = torch.randn(hidden, generator=g) * 0.1
            p[f"{prefix}.{blk}.main_branch.{b}.running_mean"] = torch.randn(hidden, generator=g) * 0.1
            p[f"{prefix}.{blk}.main_branch.{b}.running_var"] = torch.rand(hidden, generator=g) * 0.5 + 0.25
    bound = 1.0 / (hidden * upsample * upsample) ** 0.5
    p[f"{prefix}.4.weight"] = (torch.rand(hidden, hidden, upsample, upsample, generator=g) * 2 - 1) * bound * 3
    p[f"{prefix}.4.bias"] = (torch.rand(hidden, generator=g) * 2 - 1) * bound
    p[f"{prefix}.7.weight"], p[f"{prefix}.7.bias"] = conv(3, hidden, 1)
    return p
