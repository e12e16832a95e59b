"""TEST INFRASTRUCTURE ONLY -- pin the training-mode actor flip of NeuRADHashEncoding to the reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_train_encoding

Builds the unmodified reference model for tests/golden/nff_actors.npz, puts `model.field.hashgrid` in `.train()` mode and
calls its forward on the gaussians of the reference's own final samples.  The module draws one flip per ray with
``torch.bernoulli`` (field_components/neurad_encoding.py:212-219); re-seeding the global generator and drawing the same
tensor afterwards recovers the flips, which are then given to the oracle (neurad_oracle.hashgrid_forward(flip=)).  Asserts
bit-equality of features and directions and writes tests/golden/train_encoding.npz.
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


def main():
    z = np.load(os.path.join(GOLDEN, "nff_actors.npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["__meta__"]))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    rays = {k[len("ray/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ray/")}
    starts, ends = torch.from_numpy(z["ref/starts"]), torch.from_numpy(z["ref/ends"])
    cfg = nsb.small_config(n_actors=meta["n_actors"], log2_main=meta["log2_main"], log2_prop=meta["log2_prop"],
                           static_scale=meta["static_scale"], duration=meta["duration"], num_sensors=meta["num_sensors"])
    ocfg = to_oracle_cfg(cfg)
    trajs = scene.make_trajectories(meta["n_actors"], cfg.duration, seed=meta["seed"])
    model = ref_driver.build_reference_model(cfg, params, trajs)
    from nerfstudio.utils.math import GaussiansStd

    n, s = starts.shape
    lidar = rays["is_lidar"].reshape(-1).bool()
    area = rays["pixel_area"].reshape(-1) * torch.where(lidar, 1.0, float(cfg.rgb_upsample_factor**2))
    o, d = rays["origins"], rays["directions"]
    mean, std = O.fast_isotropic_gaussian(o[:, None, :], d[:, None, :], area[:, None, None], starts[..., None], ends[..., None])
    times = rays["times"].reshape(n, 1, 1).expand(n, s, 1)
    dirs = d[:, None, :].expand(n, s, 3)
    enc = model.field.hashgrid
    enc.train()
    with torch.no_grad():
        ef, ed = O.hashgrid_forward(params, "field", ocfg.main, ocfg, mean, std, times, dirs)
    for seed in range(321, 400):  # few rays cross an actor in this case: take the first seed that flips some of them
        torch.manual_seed(seed)
        with torch.no_grad():
            feats, dirs_out = enc(GaussiansStd(mean, std), times, dirs)
        torch.manual_seed(seed)
        flip = torch.bernoulli(torch.full((n,), float(enc.config.actor.flip_prob))) * -2 + 1
        with torch.no_grad():
            of, od = O.hashgrid_forward(params, "field", ocfg.main, ocfg, mean, std, times, dirs, flip=flip)
        assert torch.equal(feats, of) and torch.equal(dirs_out, od), "oracle(flip) != reference in train mode"
        changed = int((of != ef).any(-1).sum())
        kept = int(((of == ef).all(-1).view(n, s) & (O_actor_rows := (ed != dirs).any(-1))).sum())
        if changed >= 3 and kept >= 1 and not torch.equal(od, ed):
            break
    else:
        raise AssertionError("no seed flipped a ray with actor samples")
    print(f"train-mode encoding: oracle == reference bit-for-bit; {int((flip < 0).sum())}/{n} rays flipped, {changed} feature rows differ from eval mode")
    out = {"in/flip": flip.numpy(), "ref/features": feats.numpy(), "ref/directions": dirs_out.numpy(),
           "__meta__": np.array(repr(dict(case="nff_actors.npz", seed=seed, torch=torch.__version__)))}
    path = os.path.join(GOLDEN, "train_encoding.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
