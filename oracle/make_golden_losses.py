"""TEST INFRASTRUCTURE ONLY -- pin the loss restatements (oracle/losses_oracle.py) to the real reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden_losses

Inputs are the `weights_list` / `ray_samples_list` of a real render: the oracle's trace on the first rays of
tests/golden/nff_actors.npz (bit-identical to the reference's, oracle/make_golden.py), sliced like
NeuRADModel.get_nff_outputs does (the sky sample is dropped from the final level, neurad.py:385-386).  The reference's own
``zipnerf_interlevel_loss`` / ``distortion_loss`` (nerfstudio/model_components/losses.py) run on RaySamples stand-ins that
expose ``spacing_starts`` / ``spacing_ends``; losses and their autograd gradients are asserted equal to the restatement's
and written to tests/golden/losses.npz.
"""
from __future__ import annotations

import ast
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import neurad_studio_b200 as nsb  # noqa: E402
from oracle import losses_oracle as LO  # noqa: E402
from oracle import neurad_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle.convert import to_oracle_cfg  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
N_RAYS = 128


def main():
    ref_import.install()
    from nerfstudio.model_components import losses as RL

    z = np.load(os.path.join(GOLDEN, "nff_actors.npz"), allow_pickle=False)
    meta = ast.literal_eval(str(z["__meta__"]))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    rays = {k[len("ray/"):]: torch.from_numpy(z[k])[:N_RAYS] for k in z.files if k.startswith("ray/")}
    cfg = nsb.small_config(n_actors=meta["n_actors"], log2_main=meta["log2_main"], log2_prop=meta["log2_prop"],
                           static_scale=meta["static_scale"], duration=meta["duration"], num_sensors=meta["num_sensors"])
    with torch.no_grad():
        out = O.nff_outputs(params, to_oracle_cfg(cfg), rays["origins"], rays["directions"], rays["pixel_area"], rays["times"],
                            rays["sensor_idx"], rays["is_lidar"], want_trace=True)
    tr = out["trace"]
    sdist = [tr["bins_s_0"].contiguous(), tr["bins_s_1"].contiguous(), tr["bins_s_2"][:, :-1].contiguous()]
    weights = [tr["prop_weights_0"], tr["prop_weights_1"], tr["weights"][:, :-1].contiguous()]

    def leafs():
        return [w.clone().requires_grad_(True) for w in weights]

    def stand_in(sd):
        return SimpleNamespace(spacing_starts=sd[:, :-1, None], spacing_ends=sd[:, 1:, None])

    rs_list = [stand_in(s) for s in sdist]
    wr = leafs()
    li = RL.zipnerf_interlevel_loss([w[..., None] for w in wr], rs_list)
    ld = RL.distortion_loss([w[..., None] for w in wr], rs_list)
    (li * 3 + ld * 5).backward()
    wo = leafs()
    oi = LO.zipnerf_interlevel_loss(sdist, wo)
    od = LO.distortion_loss(sdist[-1], wo[-1])
    (oi * 3 + od * 5).backward()
    assert torch.equal(li, oi) and torch.equal(ld, od), (li, oi, ld, od)
    for a, b in zip(wr, wo):
        assert torch.equal(a.grad, b.grad)
    print(f"interlevel {li.item():.6e} distortion {ld.item():.6e}: restatement == reference bit-for-bit (values and gradients)")
    arrays = {}
    for i in range(3):
        arrays[f"in/sdist_{i}"] = sdist[i].numpy()
        arrays[f"in/weights_{i}"] = weights[i].numpy()
        arrays[f"ref/grad_{i}"] = wr[i].grad.numpy()
    arrays["ref/interlevel"] = li.detach().numpy()
    arrays["ref/distortion"] = ld.detach().numpy()
    # ---- lidar carving masks: the reference's NeuRADModel._compute_is_close_to_lidar (neurad.py:677-700) on stand-ins
    import nerfstudio.models.neurad as ref_neurad

    gen = torch.Generator().manual_seed(9)
    n, s = 64, 32
    edges = torch.cumsum(torch.rand(n, s + 1, generator=gen) * 8, dim=1)
    is_lidar = torch.rand(n, generator=gen) < 0.6
    did_return = torch.rand(n, generator=gen) < 0.7
    mid = (edges[:, :-1] + edges[:, 1:]) * 0.5
    dnorm = mid[torch.arange(n), torch.randint(0, s, (n,), generator=gen)] + (torch.rand(n, generator=gen) - 0.5) * 0.3
    loss_cfg = SimpleNamespace(carving_epsilon=0.1, non_return_lidar_distance=150.0)
    fake_self = SimpleNamespace(config=SimpleNamespace(loss=loss_cfg))
    for tag, with_return in (("with_return", True), ("no_return_key", False)):
        md = {"is_lidar": is_lidar[:, None, None].expand(n, s, 1).clone(), "directions_norm": dnorm[:, None, None].expand(n, s, 1).clone()}
        if with_return:
            md["did_return"] = did_return[:, None, None].expand(n, s, 1).clone()
        rs = SimpleNamespace(metadata=md, frustums=SimpleNamespace(starts=edges[:, :-1, None], ends=edges[:, 1:, None]))
        ref_neurad.NeuRADModel._compute_is_close_to_lidar(fake_self, rs)
        want = md["is_close_to_lidar"][..., 0]
        got = LO.is_close_to_lidar(edges, is_lidar, dnorm, did_return if with_return else None)
        assert torch.equal(want, got), tag
        arrays[f"ref/carving_{tag}"] = want.numpy()
    arrays.update({"in/carv_edges": edges.numpy(), "in/carv_is_lidar": is_lidar.numpy(), "in/carv_did_return": did_return.numpy(),
                   "in/carv_directions_norm": dnorm.numpy()})
    print("carving masks: restatement == reference (with and without did_return)")
    arrays["__meta__"] = np.array(repr(dict(case="nff_actors.npz", n_rays=N_RAYS, loss="3 * interlevel + 5 * distortion",
                                            pulse_widths=LO.PULSE_WIDTHS, torch=torch.__version__)))
    path = os.path.join(GOLDEN, "losses.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
