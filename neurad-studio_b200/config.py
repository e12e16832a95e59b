"""Configuration of the NeuRAD neural-feature-field forward path.

Mirrors the subset of the reference's config tree that shapes ``NeuRADModel.get_nff_outputs``:
``NeuRADModelConfig`` / ``SamplingSettings`` (nerfstudio/models/neurad.py:97-162), ``NeuRADFieldConfig`` /
``NeuRADProposalFieldConfig`` (nerfstudio/fields/neurad_field.py:44-75, 155-182) and ``StaticSettings`` /
``ActorSettings`` (nerfstudio/field_components/neurad_encoding.py:34-66).  Field names follow the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Tuple

import numpy as np
import torch


@dataclass
class HashGridSettings:
    """One multi-resolution hash grid (HashEncoding, field_components/encodings.py:326-352)."""

    hashgrid_dim: int = 4  # features per level
    num_levels: int = 8
    base_res: int = 32
    max_res: int = 8192
    log2_hashmap_size: int = 22

    @property
    def hash_table_size(self) -> int:
        return 2**self.log2_hashmap_size

    @property
    def out_dim(self) -> int:
        return self.num_levels * self.hashgrid_dim

    def scalings(self) -> torch.Tensor:
        """Per-level resolutions; the same expression as encodings.py:348-350 so that the buffer is identical
        (main grid: [32, 70, 156, 344, 760, 1680, 3709, 8191] -- note 8191)."""
        levels = torch.arange(self.num_levels)
        growth = (
            np.exp((np.log(self.max_res) - np.log(self.base_res)) / (self.num_levels - 1))
            if self.num_levels > 1
            else 1.0
        )
        return torch.floor(self.base_res * growth**levels)


def _main_static() -> HashGridSettings:
    return HashGridSettings(4, 8, 32, 8192, 22)


def _main_actor() -> HashGridSettings:
    return HashGridSettings(4, 4, 64, 1024, 17)


def _prop_static() -> HashGridSettings:
    return HashGridSettings(1, 6, 128, 4096, 20)


def _prop_actor() -> HashGridSettings:
    return HashGridSettings(1, 4, 64, 1024, 15)


@dataclass
class NeuRADHashEncodingConfig:
    static: HashGridSettings = field(default_factory=_main_static)
    actor: HashGridSettings = field(default_factory=_main_actor)
    actor_scale: float = 10.0
    # ActorSettings.flip_prob, training mode only: 0.25 for the main field's grid (fields/neurad_field.py:51), the
    # ActorSettings default 0.5 for the proposal fields' grids (field_components/neurad_encoding.py:50)
    flip_prob: float = 0.25


def _prop_grid() -> NeuRADHashEncodingConfig:
    return NeuRADHashEncodingConfig(static=_prop_static(), actor=_prop_actor(), flip_prob=0.5)


@dataclass
class SamplingSettings:
    num_proposal_samples: Tuple[int, int] = (128, 64)
    num_nerf_samples: int = 32
    power_lambda: float = -1.0
    power_scaling: float = 0.1
    sky_distance: float = 20000.0
    histogram_padding: float = 0.01  # PDFSampler default, ray_samplers.py:272
    single_jitter: bool = True  # SamplingSettings.single_jitter (neurad.py:101), training mode only


@dataclass
class NeuRADConfig:
    """Everything the forward path needs to know that is not a learned tensor."""

    grid: NeuRADHashEncodingConfig = field(default_factory=NeuRADHashEncodingConfig)
    proposal_grid_1: NeuRADHashEncodingConfig = field(default_factory=_prop_grid)
    proposal_grid_2: NeuRADHashEncodingConfig = field(default_factory=_prop_grid)
    sampling: SamplingSettings = field(default_factory=SamplingSettings)
    geo_hidden_dim: int = 32
    nff_hidden_dim: int = 32
    nff_out_dim: int = 32
    num_multisamples: int = 1  # NeuRADFieldConfig.num_multisamples, neurad_field.py:67
    appearance_dim: int = 16
    temporal_appearance_freq: float = 1.0
    rgb_upsample_factor: int = 3
    rgb_hidden_dim: int = 32
    actor_bbox_padding: Tuple[float, float, float] = (0.25, 0.25, 0.1)
    carving_epsilon: float = 0.1  # LossSettings (neurad.py:79,87): lidar carving masks of the training outputs
    non_return_lidar_distance: float = 150.0
    # scene-level constants (dataset metadata in the reference)
    static_scale: float = 100.0
    duration: float = 8.0
    num_sensors: int = 7
    n_actors: int = 0

    @property
    def proposal_grids(self):
        return (self.proposal_grid_1, self.proposal_grid_2)

    @property
    def embeds_per_sensor(self) -> int:
        return math.ceil(self.duration * self.temporal_appearance_freq)

    @property
    def feature_dim(self) -> int:
        return self.nff_out_dim + self.appearance_dim


def small_config(n_actors: int = 0, log2_main: int = 12, log2_prop: int = 11, **kw) -> NeuRADConfig:
    """A shrunken-table configuration (same levels / resolutions / code path, fewer hash slots) used for
    self-contained golden fixtures and smoke tests."""
    cfg = NeuRADConfig(n_actors=n_actors, **kw)
    cfg.grid.static.log2_hashmap_size = log2_main
    cfg.grid.actor.log2_hashmap_size = max(log2_main - 3, 6)
    for g in cfg.proposal_grids:
        g.static.log2_hashmap_size = log2_prop
        g.actor.log2_hashmap_size = max(log2_prop - 3, 6)
    return cfg
