"""TEST INFRASTRUCTURE ONLY -- drive the UNMODIFIED reference (imported from /root/reference) on given inputs.

Builds the real ``NeuRADModel`` (nerfstudio/models/neurad.py:165) with ``implementation="torch"`` on CPU, loads
a parameter dict keyed by the reference's own state_dict names, and runs ``get_nff_outputs`` while recording
the intermediate tensors the parity tests pin.  Only usable in the build container (needs /root/reference).

Two documented deviations from "unmodified", both forced by the environment:
  * ``nerfacc`` is not installed -> ``render_weight_from_alpha`` / ``accumulate_along_rays`` are provided from
    their published dense-tensor definitions (see oracle/neurad_oracle.py header);
  * the reference's CPU debugging branch in ``_render_weights`` (neurad.py:713-715, constant 0.5 weights) is
    bypassed by calling the CUDA branch (:716-717) on CPU tensors.
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional

import torch

from . import ref_import


def _install_nerfacc_restatements():
    import nerfacc  # the stub module

    def render_weight_from_alpha(alphas, packed_info=None, ray_indices=None, n_rays=None, prefix_trans=None):
        trans = torch.cumprod(torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas[..., :-1]], dim=-1), dim=-1)
        return alphas * trans, trans

    def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
        assert ray_indices is None
        if values is None:
            return torch.sum(weights[..., None], dim=-2)
        return torch.sum(weights[..., None] * values, dim=-2)

    nerfacc.render_weight_from_alpha = render_weight_from_alpha
    nerfacc.accumulate_along_rays = accumulate_along_rays


def build_reference_model(cfg, params: Dict[str, torch.Tensor], trajectories: Optional[List[dict]]):
    """cfg: neurad_studio_b200.NeuRADConfig (used only for its numbers)."""
    ref_import.install()
    _install_nerfacc_restatements()
    warnings.filterwarnings("ignore")
    import nerfstudio.models.neurad as ref_neurad
    from nerfstudio.data.scene_box import SceneBox
    from nerfstudio.field_components.field_heads import FieldHeadNames

    ref_neurad.VGGPerceptualLossPix2Pix = lambda: torch.nn.Identity()
    mc = ref_neurad.NeuRADModelConfig(implementation="torch")

    def copy_grid(dst, src):
        for name in ("static", "actor"):
            d, s = getattr(dst, name), getattr(src, name)
            d.hashgrid_dim, d.num_levels, d.base_res, d.max_res = s.hashgrid_dim, s.num_levels, s.base_res, s.max_res
            d.log2_hashmap_size = s.log2_hashmap_size
        dst.actor.actor_scale = src.actor_scale

    copy_grid(mc.field.grid, cfg.grid)
    copy_grid(mc.sampling.proposal_field_1.grid, cfg.proposal_grid_1)
    copy_grid(mc.sampling.proposal_field_2.grid, cfg.proposal_grid_2)
    mc.sampling.num_proposal_samples = tuple(cfg.sampling.num_proposal_samples)
    mc.sampling.num_nerf_samples = cfg.sampling.num_nerf_samples
    s = float(cfg.static_scale)
    scene_box = SceneBox(aabb=torch.tensor([[-s, -s, -10.0], [s, s, 30.0]]))
    metadata = {
        "duration": cfg.duration,
        "sensor_idx_to_name": {i: f"sensor{i}" for i in range(cfg.num_sensors)},
        "trajectories": trajectories if trajectories is not None else [],
    }
    model = ref_neurad.NeuRADModel(mc, scene_box=scene_box, num_train_data=1, metadata=metadata)
    sd = {k: v for k, v in params.items() if k != "static_scale"}
    res = model.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    ours_missing = [
        k
        for k in res.missing_keys
        if k.split(".")[0] in ("field", "proposal_fields", "lidar_decoder", "appearance_embedding")
        and ".hashgrid.actors." not in k  # aliases of the shared `dynamic_actors` module, loaded through that name
        and not k.endswith("beta_min")
    ]
    assert not ours_missing, ours_missing
    model.eval()

    def _render_weights(self, outputs, ray_samples):  # neurad.py:716-717 (the non-CPU branch)
        import nerfacc

        value = outputs[FieldHeadNames.ALPHA].squeeze(-1)
        weights, _ = nerfacc.render_weight_from_alpha(value)
        return weights

    model._render_weights = _render_weights.__get__(model)
    return model


def run_reference_nff(model, rays: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Run get_nff_outputs (neurad.py:368-421) and capture intermediates."""
    from nerfstudio.cameras.rays import RayBundle
    from nerfstudio.field_components.field_heads import FieldHeadNames

    n = rays["origins"].shape[0]
    rb = RayBundle(
        origins=rays["origins"].clone(),
        directions=rays["directions"].clone(),
        pixel_area=rays["pixel_area"].clone(),
        fars=rays["fars"].clone() if "fars" in rays else torch.full((n, 1), 1_000_000.0),
        times=rays["times"].clone(),
        metadata={"is_lidar": rays["is_lidar"].clone(), "sensor_idxs": rays["sensor_idx"].clone()},
    )
    cap: Dict[str, torch.Tensor] = {}

    # capture sampler outputs
    orig_sampler_forward = model.sampler.forward

    def sampler_forward(*a, **k):
        rs, wl, rsl = orig_sampler_forward(*a, **k)
        for i, (w, r) in enumerate(zip(wl, rsl)):
            cap[f"prop_weights_{i}"] = w[..., 0].clone()
            cap[f"bins_e_{i}"] = torch.cat([r.frustums.starts[..., 0], r.frustums.ends[..., -1:, 0]], -1).clone()
            cap[f"bins_s_{i}"] = torch.cat([r.spacing_starts[..., 0], r.spacing_ends[..., -1:, 0]], -1).clone()
        i = len(wl)
        cap[f"bins_e_{i}"] = torch.cat([rs.frustums.starts[..., 0], rs.frustums.ends[..., -1:, 0]], -1).clone()
        cap[f"bins_s_{i}"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1).clone()
        return rs, wl, rsl

    model.sampler.forward = sampler_forward
    orig_field_forward = model.field.forward

    def field_forward(ray_samples, *a, **k):
        out = orig_field_forward(ray_samples, *a, **k)
        cap["sdf"] = out[FieldHeadNames.SDF][..., 0].clone()
        cap["alpha"] = out[FieldHeadNames.ALPHA][..., 0].clone()
        cap["field_feature"] = out[FieldHeadNames.FEATURE].clone()
        cap["starts"] = ray_samples.frustums.starts[..., 0].clone()
        cap["ends"] = ray_samples.frustums.ends[..., 0].clone()
        return out

    model.field.forward = field_forward
    try:
        with torch.no_grad():
            out = model.get_nff_outputs(rb, calc_lidar_losses=False)
    finally:
        model.sampler.forward = orig_sampler_forward
        model.field.forward = orig_field_forward
    res = {k: v.clone() for k, v in out.items() if isinstance(v, torch.Tensor)}
    res.update(cap)
    with torch.no_grad():
        o = model.lidar_decoder(out["features"])
        res["intensity"] = o[..., 0:1].sigmoid()
        res["ray_drop_logits"] = o[..., 1:2]
    return res
