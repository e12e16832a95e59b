"""TEST SCAFFOLDING ONLY -- ctypes driver for the host SIMT emulation of the render kernel (emul.cpp)."""
import ctypes
import os
import subprocess

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libnffemul.so")
SRC = os.path.join(HERE, "emul.cpp")
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "neurad-studio_b200", "csrc")


def build(force=False):
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("nff_device.h", "nff_lane.h", "nff_params.h", "simt.h")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(
            ["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-o", SO, SRC]
        )
    return SO


def render(cfg, params, rays, pdf_u, field_of_round=(2, 2), lane_mode=False):
    """cfg: neurad_studio_b200.NeuRADConfig; params: reference-named tensors (CPU); rays: dict of CPU tensors."""
    lib = ctypes.CDLL(build())
    lib.emul_render.restype = ctypes.c_int
    keep, ptrs, ints, floats = [], [], [], []

    def P(t):
        if t is None:
            ptrs.append(None)
            return
        t = t.contiguous()
        keep.append(t)
        ptrs.append(ctypes.c_void_p(t.data_ptr()))

    n_actors = cfg.n_actors
    n_times = int(params["dynamic_actors.unique_timestamps"].shape[0]) if n_actors else 0
    ints += [n_actors, n_times]
    prefixes = ["field", "proposal_fields.0", "proposal_fields.1"]
    gcfgs = [cfg.grid, cfg.proposal_grid_1, cfg.proposal_grid_2]
    for pre, g in zip(prefixes, gcfgs):
        for name, s in (("static_grid", g.static), ("actor_grids.0", g.actor)):
            ints += [s.num_levels, s.hashgrid_dim, s.log2_hashmap_size]
            key = f"{pre}.hashgrid.{name}.scalings"
            sc = params[key].tolist() if key in params else s.scalings().tolist()
            floats += sc + [0.0] * (16 - len(sc))
        P(params[f"{pre}.hashgrid.static_grid.hash_table"])
        if n_actors:
            tabs = [params[f"{pre}.hashgrid.actor_grids.{a}.hash_table"].contiguous() for a in range(n_actors)]
            keep.extend(tabs)
            arr = (ctypes.c_void_p * n_actors)(*[t.data_ptr() for t in tabs])
            keep.append(arr)
            ptrs.append(ctypes.cast(arr, ctypes.c_void_p))
        else:
            ptrs.append(None)
        P(params.get(f"{pre}.density_decoder.weight"))
        floats += [float(params["static_scale"]), g.actor_scale]
    for k in ("field.mlp_geo.layers.0", "field.mlp_geo.layers.1", "field.mlp_feature.layers.0",
              "field.mlp_feature.layers.1", "field.mlp_feature.layers.2"):
        P(params[k + ".weight"])
        P(params[k + ".bias"])
    floats.append(float(params["field.sdf_to_density.beta"].abs() + 0.0001))
    if n_actors:
        P(params["dynamic_actors.unique_timestamps"])
        P(params["dynamic_actors.actor_rotations_6d"])
        P(params["dynamic_actors.actor_positions"])
        P(params["dynamic_actors.actor_present_at_time"].to(torch.uint8))
        P(params["dynamic_actors.actor_sizes"])
    else:
        for _ in range(5):
            ptrs.append(None)
    floats += list(cfg.actor_bbox_padding)
    sp = cfg.sampling
    floats += [sp.power_lambda, sp.power_scaling, sp.sky_distance, sp.histogram_padding, float(cfg.rgb_upsample_factor**2)]
    P(pdf_u(sp.num_proposal_samples[1]))
    P(pdf_u(sp.num_nerf_samples))
    ints += list(field_of_round)
    P(params["appearance_embedding.weight"])
    ints += [int(params["appearance_embedding.weight"].shape[0]), cfg.appearance_dim, cfg.embeds_per_sensor]
    floats.append(cfg.duration)
    n = rays["origins"].shape[0]
    P(rays["origins"].float())
    P(rays["directions"].float())
    P(rays["pixel_area"].reshape(-1).float())
    P(rays["times"].reshape(-1).float())
    P(rays["nears"].reshape(-1).float() if "nears" in rays else None)
    P(rays["fars"].reshape(-1).float() if "fars" in rays else None)
    P(rays["sensor_idx"].reshape(-1).long() if "sensor_idx" in rays else None)
    P(rays["is_lidar"].reshape(-1).to(torch.uint8) if "is_lidar" in rays else None)
    S0, S1 = sp.num_proposal_samples
    S2 = sp.num_nerf_samples
    out = {
        "features": torch.zeros(n, cfg.feature_dim), "depth": torch.zeros(n, 1), "accumulation": torch.zeros(n, 1),
        "prop_depth_0": torch.zeros(n, 1), "prop_depth_1": torch.zeros(n, 1),
        "prop_weights_0": torch.zeros(n, S0), "prop_weights_1": torch.zeros(n, S1),
        "bins_s_1": torch.zeros(n, S1 + 1), "bins_e_1": torch.zeros(n, S1 + 1),
        "bins_s_2": torch.zeros(n, S2 + 1), "bins_e_2": torch.zeros(n, S2 + 1),
        "inds_1": torch.zeros(n, S1 + 1, dtype=torch.int32), "inds_2": torch.zeros(n, S2 + 1, dtype=torch.int32),
        "sdf": torch.zeros(n, S2), "alpha": torch.zeros(n, S2), "field_feature": torch.zeros(n, S2, cfg.nff_out_dim),
        "weights": torch.zeros(n, S2),
        "actor_id_0": torch.zeros(n, S0, dtype=torch.int32), "actor_id_1": torch.zeros(n, S1, dtype=torch.int32),
        "actor_id_main": torch.zeros(n, S2, dtype=torch.int32),
    }
    for k in out:
        P(out[k])
        out[k] = keep[-1]
    c_ptrs = (ctypes.c_void_p * len(ptrs))(*ptrs)
    c_ints = (ctypes.c_int * len(ints))(*ints)
    c_floats = (ctypes.c_float * len(floats))(*floats)
    rc = lib.emul_render(c_ptrs, c_ints, c_floats, ctypes.c_longlong(n), ctypes.c_int(1 if lane_mode else 0))
    assert rc == 0
    return out
