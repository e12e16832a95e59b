"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the tiny-cuda-nn operators NeuRAD uses with `implementation="tcnn"`.

PARITY UNPINNED.  tiny-cuda-nn (an unpinned `git+https://github.com/NVlabs/tiny-cuda-nn` dependency of the reference,
`Dockerfile:65`) is absent from this image and from the GPU box, and the reference's tests hold no vector for it
(SURVEY.md 8c), so nothing here can be checked against the real library.  Every function restates the PUBLISHED algorithm
[from memory of tiny-cuda-nn's include/tiny-cuda-nn/encodings/grid.h, spherical_harmonics.h, networks/fully_fused_mlp.cu and
bindings/torch/tinycudann/modules.py]; each restatement names the call site in the reference that fixes the configuration.
What the tests can and do pin: (i) the CUDA path against THIS restatement, (ii) properties that do not depend on the
restatement being right in every detail (interpolation of a grid filled from a trilinear function is exact, dense levels
index linearly, hashed levels collide exactly where the hash says), (iii) agreement with the torch-layout operators where the
two semantics coincide.

Deliberate difference from the real library, stated once: tiny-cuda-nn evaluates in fp16 (parameters cast to half at
forward time, half accumulators in the grid kernel, half activations between MLP layers).  Here -- and in the CUDA path --
the PARAMETERS are rounded to fp16 (that is what a tcnn-trained checkpoint means) but the arithmetic is fp32.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

PRIMES = (1, 2654435761, 805459861, 3674653429)  # coherent_prime_hash (grid.h); the first three are the torch path's too


@dataclass
class TcnnGridLayout:
    """Per-level constants of a tcnn `HashGrid` encoding (encoding_config built at field_components/encodings.py:386-401:
    n_levels, n_features_per_level, log2_hashmap_size, base_resolution = min_res, per_level_scale = growth_factor)."""

    n_dims: int
    n_levels: int
    n_features: int
    scale: List[float]       # grid_scale(level) = exp2(level * log2(per_level_scale)) * base_resolution - 1   (fp32)
    resolution: List[int]    # ceil(scale) + 1 grid VERTICES per axis
    size: List[int]          # entries of the level: min(next_multiple(resolution^n_dims, 8), 2^log2_hashmap_size)
    offset: List[int]        # first entry of the level in the flat parameter vector (entries, not floats)
    dense: List[bool]        # resolution^n_dims fits the level: linear indexing, no hash

    @property
    def n_params(self) -> int:
        return (self.offset[-1] + self.size[-1]) * self.n_features


def grid_layout(n_levels: int, n_features: int, log2_hashmap_size: int, base_resolution: int, per_level_scale: float,
                n_dims: int = 3) -> TcnnGridLayout:
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    scale, res, size, offset, dense = [], [], [], [], []
    off = 0
    for lvl in range(n_levels):
        s = np.float32(np.exp2(np.float32(lvl) * log2_pls)) * np.float32(base_resolution) - np.float32(1.0)
        r = int(math.ceil(float(s))) + 1
        n = r**n_dims
        n = min(n, (2**32 - 1) // 2)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        scale.append(float(s))
        res.append(r)
        size.append(n)
        offset.append(off)
        # grid_index(): the linear index is kept when the strides never exceed the level's size
        stride, is_dense = 1, True
        for _ in range(n_dims):
            if stride > n:
                break
            stride *= r
        is_dense = not (n < stride)
        dense.append(is_dense)
        off += n
    return TcnnGridLayout(n_dims, n_levels, n_features, scale, res, size, offset, dense)


def grid_index(layout: TcnnGridLayout, lvl: int, pos_grid: Tensor) -> Tensor:
    """grid_index<N_DIMS>(): pos_grid [..., n_dims] int64 -> entry index inside the level."""
    n, r = layout.size[lvl], layout.resolution[lvl]
    if layout.dense[lvl]:
        idx = torch.zeros_like(pos_grid[..., 0])
        stride = 1
        for d in range(layout.n_dims):
            idx = idx + pos_grid[..., d] * stride
            stride *= r
    else:
        idx = torch.zeros_like(pos_grid[..., 0])
        for d in range(layout.n_dims):
            idx = idx ^ ((pos_grid[..., d] * PRIMES[d]) & 0xFFFFFFFF)
    return idx % n


def half_round(params: Tensor) -> Tensor:
    """The fp32 master parameters of the torch binding, as the kernels see them: cast to half at forward time."""
    return params.to(torch.float16).to(torch.float32)


def hashgrid_encode(layout: TcnnGridLayout, params: Tensor, x: Tensor) -> Tensor:
    """tcnn::GridEncoding forward (kernel_grid), linear interpolation.  params: flat [n_params] (fp16-representable
    values), x [P, n_dims] in [0,1] -> [P, n_levels * n_features] (level-major, like the torch binding's output)."""
    F_ = layout.n_features
    table = params.reshape(-1, F_)
    outs = []
    for lvl in range(layout.n_levels):
        # pos_fract(): fmaf(scale, x, 0.5f) -- one rounding (the product and the sum are exact in double)
        pos = (x.double() * float(np.float32(layout.scale[lvl])) + 0.5).float()
        base = torch.floor(pos)
        frac = pos - base
        base = base.to(torch.int64)
        acc = torch.zeros(x.shape[0], F_)
        for corner in range(1 << layout.n_dims):
            w = torch.ones(x.shape[0])
            pg = base.clone()
            for d in range(layout.n_dims):
                if corner & (1 << d):
                    w = w * frac[:, d]
                    pg[:, d] += 1
                else:
                    w = w * (1 - frac[:, d])
            idx = grid_index(layout, lvl, pg) + layout.offset[lvl]
            acc = acc + w[:, None] * table[idx]
        outs.append(acc)
    return torch.cat(outs, dim=-1)


def sh4(directions: Tensor) -> Tensor:
    """tcnn SphericalHarmonics, degree 4 (spherical_harmonics.h).  The reference feeds `(d + 1) / 2`
    (fields/base_field.py:136-142 -> field_components/encodings.py:803-805) and tcnn maps it back with `x * 2 - 1`: the
    polynomial is evaluated at the DIRECTION itself -- unlike the torch twin (`components_from_spherical_harmonics`,
    utils/math.py:31-94), which evaluates at the shifted value and uses the opposite sign on the odd-m terms."""
    x, y, z = directions[..., 0], directions[..., 1], directions[..., 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    c = torch.empty(*directions.shape[:-1], 16)
    c[..., 0] = 0.28209479177387814
    c[..., 1] = -0.48860251190291987 * y
    c[..., 2] = 0.48860251190291987 * z
    c[..., 3] = -0.48860251190291987 * x
    c[..., 4] = 1.0925484305920792 * xy
    c[..., 5] = -1.0925484305920792 * yz
    c[..., 6] = 0.94617469575755997 * z2 - 0.31539156525251999
    c[..., 7] = -1.0925484305920792 * xz
    c[..., 8] = 0.54627421529603959 * x2 - 0.54627421529603959 * y2
    c[..., 9] = 0.59004358992664352 * y * (-3.0 * x2 + y2)
    c[..., 10] = 2.8906114426405538 * xy * z
    c[..., 11] = 0.45704579946446572 * y * (1.0 - 5.0 * z2)
    c[..., 12] = 0.3731763325901154 * z * (5.0 * z2 - 3.0)
    c[..., 13] = 0.45704579946446572 * x * (1.0 - 5.0 * z2)
    c[..., 14] = 1.4453057213202769 * z * (x2 - y2)
    c[..., 15] = 0.59004358992664352 * x * (-x2 + 3.0 * y2)
    return c


def mlp_shapes(in_dim: int, n_neurons: int, n_hidden_layers: int, out_dim: int) -> List[Tuple[int, int]]:
    """FullyFusedMLP weight matrices, in the order they are laid out in `params` (row-major [out, in] each): the input
    layer [n_neurons, pad16(in)], n_hidden_layers - 1 hidden layers [n_neurons, n_neurons], the output layer
    [pad16(out), n_neurons].  No biases (network_config built at field_components/mlp.py:116-140)."""
    pad = lambda v: (v + 15) // 16 * 16  # noqa: E731
    shapes = [(n_neurons, pad(in_dim))]
    shapes += [(n_neurons, n_neurons)] * (n_hidden_layers - 1)
    shapes.append((pad(out_dim), n_neurons))
    return shapes


def mlp_unpack(params: Tensor, in_dim: int, n_neurons: int, n_hidden_layers: int, out_dim: int) -> List[Tensor]:
    """Flat tcnn network params -> nn.Linear-style weights [out_i, in_i] with the padding stripped (fp16-rounded)."""
    ws, off = [], 0
    shapes = mlp_shapes(in_dim, n_neurons, n_hidden_layers, out_dim)
    for i, (o, k) in enumerate(shapes):
        w = half_round(params[off:off + o * k]).reshape(o, k)
        off += o * k
        if i == 0:
            w = w[:, :in_dim]
        if i == len(shapes) - 1:
            w = w[:out_dim]
        ws.append(w.contiguous())
    assert off == params.numel(), (off, params.numel())
    return ws


def mlp_forward(ws: List[Tensor], x: Tensor) -> Tensor:
    """ReLU hidden activations, no output activation, no biases (fp32 arithmetic on the fp16-rounded weights)."""
    h = x
    for i, w in enumerate(ws):
        h = h @ w.t()
        if i < len(ws) - 1:
            h = torch.relu(h)
    return h
