"""The per-ray regularisers NeuRAD trains with, same names and call signatures as nerfstudio/model_components/losses.py
(`distortion_loss`, `zipnerf_interlevel_loss`, `ray_samples_to_sdist`; selected at models/neurad.py:262,524,541-545),
evaluated by the library's loss kernels (one thread per ray: blur, piecewise-quadratic cdf, resampling, and the analytic
gradient in the same pass) instead of ~40 small torch kernels with sorts and gathers per proposal level.

`weights_list` / `ray_samples_list` are the lists `NeuRADModel.get_nff_outputs` returns in the module walk: proposal
levels first, the final level (without the sky sample) last; weights [N,S,1].
"""
from __future__ import annotations

from typing import List

import torch
from torch import Tensor

from . import autograd as AG
from . import nerfstudio_api as api

PULSE_WIDTHS = (0.03, 0.003)  # losses.py:651


def ray_samples_to_sdist(ray_samples) -> Tensor:
    """losses.py:119-125: the spacing-domain bin edges [N,S+1] of a level."""
    return ray_samples.per_ray_spacing_bins()


def _w(weights: Tensor) -> Tensor:
    return (weights[..., 0] if weights.dim() == 3 else weights).contiguous()


def distortion_loss(weights_list: List[Tensor], ray_samples_list) -> Tensor:
    """losses.py:172-177 (mip-NeRF 360): mean over rays of lossfun_distortion on the final level; differentiable with
    respect to its weights."""
    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = _w(weights_list[-1])
    be = api.get_backend(w.device)
    if torch.is_grad_enabled() and w.requires_grad:
        return AG.DistortionLossFn.apply(be, c, w).mean()
    with torch.no_grad():
        return be.distortion_loss(c, w)[0].mean()


def zipnerf_interlevel_loss(weights_list: List[Tensor], ray_samples_list) -> Tensor:
    """losses.py:645-705 (Zip-NeRF's anti-aliased interlevel loss): the final level is the detached target, every
    proposal level receives a gradient through its weights."""
    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = _w(weights_list[-1]).detach()
    be = api.get_backend(w.device)
    loss = 0
    for i, (ray_samples, weights) in enumerate(zip(ray_samples_list[:-1], weights_list[:-1])):
        cp = ray_samples_to_sdist(ray_samples).detach()
        wp = _w(weights)
        if torch.is_grad_enabled() and wp.requires_grad:
            per_ray = AG.InterlevelLossFn.apply(be, c, w, cp, wp, PULSE_WIDTHS[i])
        else:
            with torch.no_grad():
                per_ray = be.zipnerf_interlevel_loss(c, w, cp, wp, PULSE_WIDTHS[i])[0]
        loss = loss + per_ray.mean()
    return loss
