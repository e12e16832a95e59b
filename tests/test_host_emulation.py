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
        assert int((out[k].long() != ref[k].long()).sum()) == 0, k  # exact on the reference goldens
    for k in ("features", "accumulation", "prop_depth_0", "prop_depth_1", "prop_weights_0"):
        assert rel_to_max(out[k], ref[k]) < 1e-4, k
    assert rel_to_max(out["depth"], ref["depth"]) < (2e-4 if meta["beta"] >= 20 else 1e-4)
    for k in ("bins_s_1", "bins_s_2"):
        assert (out[k] - ref[k]).abs().max().item() < 3e-6, k
    # per-sample traces: 1e-5 where the sample's two edges are bit-identical to the reference's, 2e-3 elsewhere
    # (an edge one ulp off moves the sample; the finest grid level amplifies that to ~1e-4 -- tests/test_reference_noise_floor.py)
    eq = out["bins_s_2"] == ref["bins_s_2"]
    same = eq[:, :-1] & eq[:, 1:]
    for k in ("sdf", "alpha", "field_feature"):
        a, b = out[k].float().reshape(ref[k].shape), ref[k].float()
        scale = b.abs().max().item()
        err = (a - b).abs().reshape(*same.shape, -1).amax(-1)
        if same.any() and not lane_mode:  # warp-per-ray emulation: exact-fp32 FFMA MLP
            assert err[same].max().item() < 1e-5 * scale, k
        assert err.max().item() < 2e-3 * scale, k
