"""tiny-cuda-nn layout support: rendering NeuRAD checkpoints trained with `implementation="tcnn"`.

Real NeuRAD checkpoints come from the reference's default backend, tiny-cuda-nn (`models/neurad.py:146`), whose parameters
are NOT the torch twins' tensors: every `HashEncoding` / `MLP` holds one flat `tcnn_encoding.params` vector
(field_components/encodings.py:386-401, field_components/mlp.py:116-140), the grids are vertex-centred with dense coarse
levels and per-level sizes, all actors share one 4-D grid (field_components/neurad_encoding.py:110-131, 270-281), the MLPs
have no biases and padded widths, and the SH basis is tiny-cuda-nn's.  This module holds the host side of that layout:

  * `grid_layout(...)`      per-level constants of a tcnn `HashGrid` (scale, resolution, offset, size, dense / hashed),
  * `grid_desc(...)`        the same as the C ABI's `b200nerf_tcnn_grid_desc`,
  * `mlp_unpack(...)`       flat `FullyFusedMLP` params -> nn.Linear-shaped weights with the padding stripped,
  * `is_tcnn_state(...)`, `split_state(...)`  what `B200Backend.load_params` uses to bind such a parameter set.

PARITY UNPINNED: tiny-cuda-nn is not installed in this environment (nor is a tcnn-trained checkpoint available), so the
layout follows the library's published algorithm [from memory of encodings/grid.h, networks/fully_fused_mlp.cu and
bindings/torch/tinycudann/modules.py]; tests compare the CUDA path with an independent CPU restatement
(oracle/tcnn_oracle.py) and with layout-independent properties.  Arithmetic is fp32 on fp16-rounded parameters (tiny-cuda-nn
computes in fp16).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from .config import HashGridSettings, NeuRADConfig
from .lib import TcnnGridDesc

TCNN_SUFFIX = "tcnn_encoding.params"


def growth_factor(g: HashGridSettings) -> float:
    """HashEncoding.growth_factor (encodings.py:347), the `per_level_scale` handed to tiny-cuda-nn."""
    return float(np.exp((np.log(g.max_res) - np.log(g.base_res)) / (g.num_levels - 1))) if g.num_levels > 1 else 1.0


def grid_layout(n_levels: int, n_features: int, log2_hashmap_size: int, base_resolution: int, per_level_scale: float,
                n_dims: int = 3) -> Dict[str, list]:
    """tcnn::GridEncoding's per-level constants: scale = exp2(l * log2(per_level_scale)) * base - 1 (fp32),
    resolution = ceil(scale) + 1, entries = min(next_multiple(resolution^n_dims, 8), 2^log2_hashmap_size), entry offsets,
    and whether grid_index() keeps the linear index (the strides never exceed the level's entry count)."""
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    out: Dict[str, list] = {"scale": [], "resolution": [], "size": [], "offset": [], "dense": []}
    off = 0
    for lvl in range(n_levels):
        s = np.float32(np.exp2(np.float32(lvl) * log2_pls)) * np.float32(base_resolution) - np.float32(1.0)
        r = int(math.ceil(float(s))) + 1
        n = min(r**n_dims, (2**32 - 1) // 2)
        n = (n + 7) // 8 * 8
        n = min(n, 1 << log2_hashmap_size)
        stride = 1
        for _ in range(n_dims):
            if stride > n:
                break
            stride *= r
        out["scale"].append(float(s))
        out["resolution"].append(r)
        out["size"].append(n)
        out["offset"].append(off)
        out["dense"].append(not (n < stride))
        off += n
    out["n_entries"] = off
    out["n_dims"], out["n_features"], out["n_levels"] = n_dims, n_features, n_levels
    return out


def layout_of(g: HashGridSettings, n_dims: int = 3) -> Dict[str, list]:
    return grid_layout(g.num_levels, g.hashgrid_dim, g.log2_hashmap_size, g.base_res, growth_factor(g), n_dims)


def grid_desc(layout: Dict[str, list], scalings: Tensor) -> TcnnGridDesc:
    d = TcnnGridDesc()
    d.num_levels, d.features_per_level, d.n_input_dims = layout["n_levels"], layout["n_features"], layout["n_dims"]
    sc = scalings.detach().cpu().float().tolist()
    for i in range(layout["n_levels"]):
        d.scale[i], d.resolution[i], d.offset[i], d.size[i] = layout["scale"][i], layout["resolution"][i], layout["offset"][i], layout["size"][i]
        d.dense[i] = 1 if layout["dense"][i] else 0
        d.scalings[i] = sc[i]
    return d


def half_round(t: Tensor) -> Tensor:
    """The torch binding keeps fp32 master parameters and casts them to half at every forward; the kernels read floats that
    hold exactly those half values."""
    return t.detach().to(torch.float16).to(torch.float32)


def mlp_shapes(in_dim: int, n_neurons: int, n_hidden_layers: int, out_dim: int) -> List[Tuple[int, int]]:
    """FullyFusedMLP weight matrices in `params` order, row-major [out, in]: input layer [n_neurons, pad16(in)],
    n_hidden_layers - 1 hidden layers, output layer [pad16(out), n_neurons]; no biases."""
    pad = lambda v: (v + 15) // 16 * 16  # noqa: E731
    return [(n_neurons, pad(in_dim))] + [(n_neurons, n_neurons)] * (n_hidden_layers - 1) + [(pad(out_dim), n_neurons)]


def mlp_unpack(params: Tensor, in_dim: int, n_neurons: int, n_hidden_layers: int, out_dim: int) -> List[Tensor]:
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
    if off != params.numel():
        raise ValueError(f"tcnn MLP parameter vector has {params.numel()} entries, the configuration needs {off}")
    return ws


def is_tcnn_state(params: Dict[str, Tensor]) -> bool:
    return f"field.hashgrid.static_grid.{TCNN_SUFFIX}" in params


def mlp_tensors(params: Dict[str, Tensor], prefix: str, in_dim: int, width: int, n_layers: int, out_dim: int, device) -> List[Tensor]:
    """[w0, b0, w1, b1, ...] (nn.Linear layout, zero biases) of one tcnn MLP of a checkpoint."""
    ws = mlp_unpack(params[f"{prefix}.{TCNN_SUFFIX}"].reshape(-1), in_dim, width, n_layers - 1, out_dim)
    out: List[Tensor] = []
    for w in ws:
        out += [w.to(device).contiguous(), torch.zeros(w.shape[0], device=device)]
    return out


def n_grid_params(cfg: NeuRADConfig) -> Dict[str, int]:
    """Expected sizes of every `tcnn_encoding.params` vector (checked when binding)."""
    sizes = {}
    for pre, g in (("field", cfg.grid), ("proposal_fields.0", cfg.proposal_grid_1), ("proposal_fields.1", cfg.proposal_grid_2)):
        ls = layout_of(g.static, 3)
        sizes[f"{pre}.hashgrid.static_grid.{TCNN_SUFFIX}"] = ls["n_entries"] * ls["n_features"]
        la = layout_of(g.actor, 4)
        sizes[f"{pre}.hashgrid.actor_grids.0.{TCNN_SUFFIX}"] = la["n_entries"] * la["n_features"]
    return sizes
