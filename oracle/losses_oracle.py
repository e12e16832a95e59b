"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the two per-ray regularisers NeuRAD trains with, on plain tensors.

Follows nerfstudio/model_components/losses.py: ``ray_samples_to_sdist`` (:119-125), ``lossfun_distortion`` /
``distortion_loss`` (:160-177), ``_blur_stepfun`` (:616-626), ``_sorted_interp_quad`` (:629-642) and
``zipnerf_interlevel_loss`` (:645-705; the model selects it at models/neurad.py:262).  Pinned bit for bit against the
imported reference functions by oracle/make_golden_losses.py (tests/golden/losses.npz).  Only tests/ may import this.

sdist: spacing-domain bin edges [N,S+1]; weights [N,S].
"""
from typing import List, Tuple

import torch
from torch import Tensor

PULSE_WIDTHS = (0.03, 0.003)  # losses.py:651


def lossfun_distortion(t: Tensor, w: Tensor) -> Tensor:
    ut = (t[..., 1:] + t[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    loss_inter = torch.sum(w * torch.sum(w[..., None, :] * dut, dim=-1), dim=-1)
    loss_intra = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return loss_inter + loss_intra


def distortion_loss(sdist: Tensor, weights: Tensor) -> Tensor:
    """mean over rays of lossfun_distortion on the final level (losses.py:172-177)."""
    return torch.mean(lossfun_distortion(sdist, weights))


def blur_stepfun(x: Tensor, y: Tensor, r: float) -> Tuple[Tensor, Tensor]:
    xr, xr_idx = torch.sort(torch.cat([x - r, x + r], dim=-1))
    y1 = (torch.cat([y, torch.zeros_like(y[..., :1])], dim=-1) - torch.cat([torch.zeros_like(y[..., :1]), y], dim=-1)) / (2 * r)
    y2 = torch.cat([y1, -y1], dim=-1).take_along_dim(xr_idx[..., :-1], dim=-1)
    yr = torch.cumsum((xr[..., 1:] - xr[..., :-1]) * torch.cumsum(y2, dim=-1), dim=-1).clamp_min(0)
    yr = torch.cat([torch.zeros_like(yr[..., :1]), yr], dim=-1)
    return xr, yr


def sorted_interp_quad(x: Tensor, xp: Tensor, fpdf: Tensor, fcdf: Tensor) -> Tensor:
    right_idx = torch.searchsorted(xp, x)
    left_idx = (right_idx - 1).clamp_min(0)
    right_idx = right_idx.clamp_max(xp.shape[-1] - 1)
    xp0 = xp.take_along_dim(left_idx, dim=-1)
    xp1 = xp.take_along_dim(right_idx, dim=-1)
    fpdf0 = fpdf.take_along_dim(left_idx, dim=-1)
    fpdf1 = fpdf.take_along_dim(right_idx, dim=-1)
    fcdf0 = fcdf.take_along_dim(left_idx, dim=-1)
    offset = torch.clip(torch.nan_to_num((x - xp0) / (xp1 - xp0), 0), 0, 1)
    return fcdf0 + (x - xp0) * (fpdf0 + fpdf1 * offset + fpdf0 * (1 - offset)) * 0.5


def zipnerf_interlevel_per_ray(c: Tensor, w: Tensor, cp: Tensor, wp: Tensor, pulse_width: float) -> Tensor:
    """One proposal level, per-ray sums (before the mean over rays of losses.py:704).  c / w: final level (detached by
    the caller like the reference does), cp / wp: proposal level."""
    accum_w = torch.sum(w, dim=-1, keepdim=True)
    w = torch.cat([w[..., :-1], w[..., -1:] + (1 - accum_w)], dim=-1)
    w_norm = w / (c[..., 1:] - c[..., :-1])
    c_, w_ = blur_stepfun(c, w_norm, pulse_width)
    area = 0.5 * (w_[..., 1:] + w_[..., :-1]) * (c_[..., 1:] - c_[..., :-1])
    cdf = torch.cat([torch.zeros_like(area[..., :1]), torch.cumsum(area, dim=-1)], dim=-1)
    c_ = torch.cat([torch.zeros_like(c_[..., :1]), c_, torch.ones_like(c_[..., :1])], dim=-1)
    w_ = torch.cat([torch.zeros_like(w_[..., :1]), w_, torch.zeros_like(w_[..., :1])], dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf, torch.ones_like(cdf[..., :1])], dim=-1)
    cdf_interp = sorted_interp_quad(cp, c_, w_, cdf)
    w_s = torch.diff(cdf_interp, dim=-1)
    return ((w_s - wp).clamp_min(0) ** 2 / (wp + 1e-5)).sum(dim=-1)


def zipnerf_interlevel_loss(sdist_list: List[Tensor], weights_list: List[Tensor]) -> Tensor:
    """losses.py:645-705 on [sdist per level], [weights per level] (the last entry is the final level)."""
    c, w = sdist_list[-1].detach(), weights_list[-1].detach()
    loss = 0
    for i, (cp, wp) in enumerate(zip(sdist_list[:-1], weights_list[:-1])):
        loss = loss + zipnerf_interlevel_per_ray(c, w, cp, wp, PULSE_WIDTHS[i]).mean()
    return loss


def is_close_to_lidar(bins_e: Tensor, is_lidar: Tensor, directions_norm: Tensor, did_return, carving_epsilon: float = 0.1,
                      non_return_lidar_distance: float = 150.0) -> Tensor:
    """NeuRADModel._compute_is_close_to_lidar (models/neurad.py:677-700) on per-ray tensors: bins_e [N,S+1] euclidean
    edges, is_lidar / did_return [N] bool, directions_norm [N] (measured distance of lidar rays) -> mask [N,S] bool."""
    sample_distance = (bins_e[:, :-1] + bins_e[:, 1:]) * 0.5
    dist = directions_norm[:, None] - sample_distance
    close_to_hit = dist.abs() < carving_epsilon
    if did_return is not None:
        in_lidar_range = sample_distance < non_return_lidar_distance
        m = (did_return[:, None] & close_to_hit) | ((~did_return[:, None]) & in_lidar_range)
    else:
        m = close_to_hit
    return m & is_lidar[:, None]
