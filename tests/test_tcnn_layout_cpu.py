"""CPU: tiny-cuda-nn layout (parity unpinned, see tests/tcnn_cases.py) -- the product's host-side layout
code against the oracle's independent restatement, and the DEVICE functions (host emulation) against the oracle."""
import pytest

from tests import tcnn_cases as C


def test_layout_has_every_kind_of_level():
    C.layout_has_every_kind_of_level()


def test_mlp_unpack_strips_the_padding():
    C.mlp_unpack_strips_the_padding()


def test_grid_matches_oracle_and_interpolates_linear_functions():
    C.grid_matches_oracle_and_interpolates_linear_functions("cpu")


@pytest.mark.parametrize("n_actors", [0, 3])
def test_emulated_fused_render_matches_tcnn_oracle(n_actors):
    C.fused_render_matches_tcnn_oracle("cpu", n_actors=n_actors, n_rays=48)


def test_mirror_model_takes_a_tcnn_checkpoint(monkeypatch):
    """NeuRADModel(implementation="tcnn") holds the reference's tcnn-mode state-dict keys: a tcnn-trained checkpoint loads with
    load_state_dict; the module walk / training refuse (inference through the fused kernels only)."""
    import torch

    from neurad_studio_b200 import scene
    from neurad_studio_b200.nerfstudio_api import NeuRADModel

    cfg = C.tcnn_test_config(2)
    trajs = scene.make_trajectories(2, cfg.duration, seed=1)
    ckpt = {"_model." + k: v for k, v in scene.make_params_tcnn(cfg, seed=8, trajectories=trajs).items() if k != "static_scale"}
    model = NeuRADModel(cfg, trajs, implementation="tcnn")
    sd = model.state_dict()
    assert "field.hashgrid.static_grid.tcnn_encoding.params" in sd and "field.mlp_geo.tcnn_encoding.params" in sd
    assert "field.hashgrid.actor_grids.0.tcnn_encoding.params" in sd and "field.hashgrid.actor_grids.1.tcnn_encoding.params" not in sd
    assert not any(k.endswith("hash_table") or ".layers." in k for k in sd if not k.startswith("rgb_decoder"))
    res = model.load_state_dict(ckpt, strict=False)
    assert all(k.startswith("rgb_decoder.") for k in res.missing_keys)
    assert torch.equal(model.state_dict()["field.mlp_feature.tcnn_encoding.params"], ckpt["_model.field.mlp_feature.tcnn_encoding.params"])
    from neurad_studio_b200.nerfstudio_api import RayBundle

    rays = scene.random_rays(4, cfg, seed=2, trajectories=trajs)
    rb = RayBundle(origins=rays["origins"], directions=rays["directions"], pixel_area=rays["pixel_area"], times=rays["times"])
    from neurad_studio_b200 import nerfstudio_api
    from tests.fake_backend import FakeBackend

    monkeypatch.setattr(nerfstudio_api, "get_backend", lambda device: FakeBackend())
    with pytest.raises(NotImplementedError):
        model.get_nff_outputs(rb, fused=False)
