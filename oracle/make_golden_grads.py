"""TEST INFRASTRUCTURE ONLY -- pin the BACKWARD pass to the real reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_grads

For the committed forward cases tests/golden/nff_static.npz and nff_actors.npz it builds the unmodified reference
``NeuRADModel`` (implementation="torch", CPU, eval-mode sampling so the run is deterministic, autograd ON), runs
``get_nff_outputs`` on the first N_RAYS rays, back-propagates the seeded linear loss

    L = sum_k <G_k, out_k>,   k in {features, depth, accumulation, prop_depth_0, prop_depth_1}

and records d L / d parameter for every parameter the path trains.  It asserts that torch autograd through the oracle
restatement gives the same gradients (the oracle's forward is bit-identical to the reference's, so its autograd graph is
the same up to summation order) and writes tests/golden/grads_<case>.npz.  tests/ then check the oracle and the
hand-written CUDA backward operators against these numbers on any machine.
"""
from __future__ import annotations

import ast
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
N_RAYS = 96
OUT_KEYS = ("features", "depth", "accumulation", "prop_depth_0", "prop_depth_1")
LOSS_SCALE = {"features": 1.0, "depth": 0.01, "accumulation": 1.0, "prop_depth_0": 0.01, "prop_depth_1": 0.01}


def loss_weights(shapes, seed=7):
    """The seeded cotangents G_k (also re-created by the tests)."""
    gen = torch.Generator().manual_seed(seed)
    return {k: torch.randn(shapes[k], generator=gen) * LOSS_SCALE[k] for k in OUT_KEYS}


POSE_KEYS = ("dynamic_actors.actor_positions", "dynamic_actors.actor_rotations_6d")  # optimize_trajectories (dynamic_actors.py:37)


def trainable_keys(params):
    return [k for k, v in params.items() if v.dtype.is_floating_point and (not k.startswith("dynamic_actors.") or k in POSE_KEYS)
            and not k.endswith("scalings") and k != "static_scale"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    meta = ast.literal_eval(str(z["__meta__"]))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    rays = {k[len("ray/"):]: torch.from_numpy(z[k])[:N_RAYS] for k in z.files if k.startswith("ray/")}
    return meta, params, rays


def oracle_grads(cfg, params, rays):
    p = {k: v.clone() for k, v in params.items()}
    keys = trainable_keys(p)
    for k in keys:
        p[k].requires_grad_(True)
    out = O.nff_outputs(p, to_oracle_cfg(cfg), rays["origins"], rays["directions"], rays["pixel_area"], rays["times"],
                        rays["sensor_idx"], rays["is_lidar"])
    G = loss_weights({k: out[k].shape for k in OUT_KEYS})
    sum((out[k] * G[k]).sum() for k in OUT_KEYS).backward()
    return {k: p[k].grad for k in keys if p[k].grad is not None}, {k: out[k].detach() for k in OUT_KEYS}


def grads_case(name):
    meta, params, rays = load_case(name)
    cfg = nsb.small_config(n_actors=meta["n_actors"], log2_main=meta["log2_main"], log2_prop=meta["log2_prop"],
                           static_scale=meta["static_scale"], duration=meta["duration"], num_sensors=meta["num_sensors"])
    trajs = scene.make_trajectories(meta["n_actors"], cfg.duration, seed=meta["seed"]) if meta["n_actors"] else None
    model = ref_driver.build_reference_model(cfg, params, trajs)  # eval mode: deterministic sampling, no actor flip
    from nerfstudio.cameras.rays import RayBundle  # importable once ref_import.install() has run

    named = dict(model.named_parameters())
    keys = [k for k in trainable_keys(params) if k in named]
    for k in keys:
        named[k].requires_grad_(True)
    n = rays["origins"].shape[0]
    rb = RayBundle(origins=rays["origins"].clone(), directions=rays["directions"].clone(), pixel_area=rays["pixel_area"].clone(),
                   fars=torch.full((n, 1), 1_000_000.0), times=rays["times"].clone(),
                   metadata={"is_lidar": rays["is_lidar"].clone(), "sensor_idxs": rays["sensor_idx"].clone()})
    out = model.get_nff_outputs(rb)
    G = loss_weights({k: out[k].shape for k in OUT_KEYS})
    sum((out[k] * G[k]).sum() for k in OUT_KEYS).backward()
    ref = {k: named[k].grad.detach().clone() for k in keys if named[k].grad is not None}
    ora, ora_out = oracle_grads(cfg, params, rays)
    for k in OUT_KEYS:
        assert torch.equal(out[k].detach(), ora_out[k]), f"oracle forward != reference for {k}"
    assert set(ref) == set(ora), (sorted(set(ref) ^ set(ora)))
    worst = 0.0
    for k in ref:
        scale = ref[k].abs().max().item()
        if scale == 0:
            assert ora[k].abs().max().item() == 0, k
            continue
        err = (ref[k] - ora[k]).abs().max().item() / scale
        worst = max(worst, err)
        assert err < 1e-5, (k, err)
    print(f"{name}: reference autograd == oracle autograd on {len(ref)} parameters (worst rel. diff {worst:.1e})")
    arrays = {f"grad/{k}": v.numpy() for k, v in ref.items()}
    arrays["__meta__"] = np.array(repr(dict(case=name, n_rays=N_RAYS, loss_seed=7, loss_scale=LOSS_SCALE, torch=torch.__version__)))
    path = os.path.join(GOLDEN, "grads_" + name)
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    for case in ("nff_static.npz", "nff_actors.npz"):
        grads_case(case)
