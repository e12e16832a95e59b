"""TEST INFRASTRUCTURE ONLY -- map the product's NeuRADConfig onto the oracle's NeuRADCfg (plain numbers)."""
from . import neurad_oracle as O


def to_oracle_cfg(cfg) -> O.NeuRADCfg:
    def g(s):
        return O.GridCfg(s.num_levels, s.base_res, s.max_res, s.log2_hashmap_size, s.hashgrid_dim)

    def f(gc):
        return O.FieldCfg(static=g(gc.static), actor=g(gc.actor), actor_scale=gc.actor_scale)

    return O.NeuRADCfg(
        main=f(cfg.grid),
        prop=(f(cfg.proposal_grid_1), f(cfg.proposal_grid_2)),
        num_proposal_samples=tuple(cfg.sampling.num_proposal_samples),
        num_nerf_samples=cfg.sampling.num_nerf_samples,
        power_lambda=cfg.sampling.power_lambda,
        power_scaling=cfg.sampling.power_scaling,
        sky_distance=cfg.sampling.sky_distance,
        histogram_padding=cfg.sampling.histogram_padding,
        appearance_dim=cfg.appearance_dim,
        temporal_appearance_freq=cfg.temporal_appearance_freq,
        rgb_upsample_factor=cfg.rgb_upsample_factor,
        nff_out_dim=cfg.nff_out_dim,
        actor_bbox_padding=tuple(cfg.actor_bbox_padding),
        static_scale=cfg.static_scale,
        duration=cfg.duration,
        num_sensors=cfg.num_sensors,
        n_actors=cfg.n_actors,
    )
