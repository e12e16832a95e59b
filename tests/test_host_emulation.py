"""CPU: the render kernel's device code (nff_device.h), executed by the host SIMT emulation in tests/host_emul,
against the reference golden vectors.  This validates the kernel logic where no GPU exists; the GPU parity tests
(tests/test_parity_gpu.py) are the real gate."""
import pytest
import torch

from oracle import neurad_oracle as O
from tests.helpers import cfg_from_meta, load_golden
from tests.host_emul import emul


def rel_to_max(a, b):
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-30)


@pytest.mark.parametrize("lane_mode", [False, True], ids=["warp_per_ray", "ray_per_lane"])
@pytest.mark.parametrize("name", ["nff_static.npz", "nff_actors.npz", "nff_sharp.npz"])
def test_emulated_kernel_matches_reference_golden(name, lane_mode):
    meta, g = load_golden(name)
    cfg = cfg_from_meta(meta)
    p, r, ref = g["param"], g["ray"], g["ref"]
    out = emul.render(cfg, p, r, O.pdf_u, lane_mode=lane_mode)
    for k in ("inds_1", "inds_2", "actor_id_0", "actor_id_1", "actor_id_main"):
        mism = (out[k].long() != ref[k].long()).float().mean().item()
        assert mism <= 1e-3, (k, mism)
    for k in ("features", "accumulation", "prop_depth_0", "prop_depth_1", "prop_weights_0"):
        assert rel_to_max(out[k], ref[k]) < 1e-4, k
    assert rel_to_max(out["depth"], ref["depth"]) < (5e-4 if meta["beta"] >= 20 else 1e-4)
    for k in ("bins_s_1", "bins_s_2"):
        assert (out[k] - ref[k]).abs().max().item() < 1e-5, k
    for k in ("sdf", "alpha", "field_feature"):
        assert rel_to_max(out[k], ref[k]) < 2e-3, k
