"""TEST INFRASTRUCTURE ONLY -- generate the committed golden vectors under tests/golden/ from the REAL reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

For every case it (1) builds the unmodified reference ``NeuRADModel`` (implementation="torch", CPU) through
oracle/ref_driver.py, (2) runs it on seeded synthetic inputs, (3) asserts that the oracle restatement
(oracle/neurad_oracle.py) reproduces the reference BIT FOR BIT on this machine, and (4) writes inputs,
parameters and reference outputs to a self-contained ``.npz`` so that the GPU box (which has no /root/reference)
can check both the oracle and the CUDA path against the reference's numbers.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import neurad_studio_b200 as nsb  # noqa: E402
from neurad_studio_b200 import scene  # noqa: E402
from oracle import neurad_oracle as O  # noqa: E402
from oracle import ref_driver  # noqa: E402
from oracle.convert import to_oracle_cfg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _save(name, arrays, meta):
    out = {}
    for k, v in arrays.items():
        v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
        out[k] = v
    out["__meta__"] = np.array(repr(meta))
    path = os.path.join(GOLDEN, name)
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1e6:.2f} MB")


def nff_case(name, n_actors, n_rays, seed, beta, sdf_bias, table_scale):
    torch.manual_seed(0)
    cfg = nsb.small_config(n_actors=n_actors, log2_main=10, log2_prop=10)
    trajs = scene.make_trajectories(n_actors, cfg.duration, seed=seed) if n_actors else None
    params = scene.make_params(cfg, seed=seed, table_scale=table_scale, beta=beta, trajectories=trajs, sdf_bias=sdf_bias)
    rays = scene.random_rays(n_rays, cfg, seed=seed + 1, trajectories=trajs)
    model = ref_driver.build_reference_model(cfg, params, trajs)

    # capture the main static grid's in/out (stage-level golden for hashgrid_fwd)
    cap = {}
    sg = model.field.hashgrid.static_grid
    orig = sg.forward

    def fwd(x):
        y = orig(x)
        cap["hash_in"], cap["hash_out"] = x.detach().clone(), y.detach().clone()
        return y

    sg.forward = fwd
    ref = ref_driver.run_reference_nff(model, rays)
    sg.forward = orig
    ref.update(cap)

    # closure late-binding check (neurad.py:248): every density fn evaluates proposal_fields[-1]
    from nerfstudio.cameras.rays import RayBundle  # noqa: F401

    ocfg = to_oracle_cfg(cfg)
    with torch.no_grad():
        out = O.nff_outputs(
            params, ocfg, rays["origins"], rays["directions"], rays["pixel_area"], rays["times"], rays["sensor_idx"],
            rays["is_lidar"], want_trace=True,
        )
        inten, drop = O.decode_lidar(params, out["features"])
    tr = out.pop("trace")
    allo = {**out, **tr, "intensity": inten, "ray_drop_logits": drop}
    for k, v in ref.items():
        if k in allo:
            assert torch.equal(v.float(), allo[k].float()), f"oracle != reference for {k}"
    assert torch.equal(ref["hash_in"].view(-1, 3), tr["static_pos_main"].view(-1, 3))
    n_hit = int((tr["actor_id_main"] >= 0).sum())
    print(f"{name}: oracle == reference bit-for-bit on {len(ref)} tensors; actor hits (main) = {n_hit}")
    # extra int goldens from the oracle trace (bit-identical pipeline, see assert above)
    ref["inds_1"], ref["inds_2"] = tr["inds_1"], tr["inds_2"]
    ref["cdf_1"], ref["cdf_2"] = tr["cdf_1"], tr["cdf_2"]
    ref["actor_id_main"], ref["actor_id_0"], ref["actor_id_1"] = tr["actor_id_main"], tr["actor_id_0"], tr["actor_id_1"]
    ref["density_0"], ref["density_1"] = tr["density_0"], tr["density_1"]
    arrays = {f"param/{k}": v for k, v in params.items()}
    arrays.update({f"ray/{k}": v for k, v in rays.items()})
    arrays.update({f"ref/{k}": v for k, v in ref.items()})
    meta = dict(n_actors=n_actors, log2_main=10, log2_prop=10, seed=seed, beta=beta, sdf_bias=sdf_bias,
                table_scale=table_scale, static_scale=cfg.static_scale, duration=cfg.duration,
                num_sensors=cfg.num_sensors, torch=torch.__version__)
    _save(name, arrays, meta)


def raygen_case():
    ref_driver.ref_import.install()
    from nerfstudio.cameras.cameras import Cameras, CameraType
    from nerfstudio.cameras.lidars import Lidars, LidarType

    cams = scene.pandaset_rig(time=3.7, width=96, height=54)
    arrays = {}
    for i in (0, 3):
        c = cams[i]
        rc = Cameras(
            camera_to_worlds=c.c2w[None], fx=c.fx, fy=c.fy, cx=c.cx, cy=c.cy, width=c.width, height=c.height,
            camera_type=CameraType.PERSPECTIVE, times=torch.tensor([c.time]),
            metadata={
                "rolling_shutter_time": torch.tensor([[c.rolling_shutter_time]]),
                "time_to_center_pixel": torch.tensor([[c.time_to_center_pixel]]),
                "velocities": c.velocity[None],
            },
        )
        rb = rc.generate_rays(camera_indices=0, keep_shape=True)
        coords = rc.get_image_coords()
        o = O.generate_rays_pinhole(c.c2w, c.fx, c.fy, c.cx, c.cy, c.height, c.width, coords, c.time, c.velocity,
                                    c.rolling_shutter_time, c.time_to_center_pixel)
        for k in ("origins", "directions", "pixel_area", "times"):
            assert torch.equal(getattr(rb, k), o[k]), f"pinhole {k}"
        assert torch.equal(rb.metadata["directions_norm"], o["directions_norm"])
        arrays.update({f"cam{i}/c2w": c.c2w, f"cam{i}/intr": torch.tensor([c.fx, c.fy, c.cx, c.cy]),
                       f"cam{i}/hw": torch.tensor([c.height, c.width]), f"cam{i}/time": torch.tensor(c.time),
                       f"cam{i}/velocity": c.velocity,
                       f"cam{i}/rs": torch.tensor([c.rolling_shutter_time, c.time_to_center_pixel]),
                       f"cam{i}/origins": rb.origins, f"cam{i}/directions": rb.directions,
                       f"cam{i}/pixel_area": rb.pixel_area, f"cam{i}/times": rb.times})
    scan = scene.pandar64_scan(time=3.7, beams=8, azimuths=90)
    rl = Lidars(lidar_to_worlds=scan.l2w[None], lidar_type=LidarType.PANDAR64, times=torch.tensor([scan.time]),
                metadata={"velocities": scan.velocity[None]})
    idx = torch.zeros_like(scan.points[:, 0:1]).long()
    rb = rl.generate_rays(lidar_indices=idx, points=scan.points, keep_shape=True)
    o = O.generate_rays_lidar_points(scan.l2w, scan.points, scan.time, scan.velocity)
    for k in ("origins", "directions", "pixel_area", "times"):
        assert torch.equal(getattr(rb, k), o[k]), f"lidar {k}"
    assert torch.equal(rb.metadata["did_return"], o["did_return"])
    arrays.update({"lidar/l2w": scan.l2w, "lidar/points": scan.points, "lidar/time": torch.tensor(scan.time),
                   "lidar/velocity": scan.velocity, "lidar/origins": rb.origins, "lidar/directions": rb.directions,
                   "lidar/pixel_area": rb.pixel_area, "lidar/times": rb.times,
                   "lidar/distance": rb.metadata["directions_norm"]})
    print("raygen: oracle == reference bit-for-bit (2 pinhole cameras with rolling shutter, 1 lidar scan)")
    _save("raygen.npz", arrays, dict(torch=torch.__version__))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    nff_case("nff_static.npz", n_actors=0, n_rays=96, seed=3, beta=3.0, sdf_bias=0.6, table_scale=1.0)
    nff_case("nff_actors.npz", n_actors=6, n_rays=96, seed=5, beta=4.0, sdf_bias=0.5, table_scale=1.0)
    nff_case("nff_sharp.npz", n_actors=0, n_rays=64, seed=7, beta=20.0, sdf_bias=None, table_scale=1.0)
    raygen_case()


if __name__ == "__main__":
    main()
